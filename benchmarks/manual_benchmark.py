#!/usr/bin/env python
"""manual_benchmark -- the reference's benchmark CLI on the sm_100a library
(SURVEY.md 8(f2)): same flags, same measurement protocol, same CSV schema as
benchmarks/manual_benchmark.cu:44-129,199-518 of the reference, so one sweep
script drives both builds on the same box.

    python benchmarks/manual_benchmark.py --num_categories 10000000 --embed_width 256 \
        --batch_size 65536 --hotness 64 --alpha=1.15 --half_embedding_type \
        --iterations 100 --enable_csv

Flags are absl-style: `--name value`, `--name=value`, booleans as `--flag`,
`--noflag` or `--flag=true|false`.  Extras beyond the reference: `--bf16`,
`--combine_mode {sum,mean,concat}`, `--csv_file`.
Timing as in the reference (:199-248): one warm-up call, then per-iteration
cudaEvent pairs with the caches cleared between iterations (`--clear_caches`).
Synthetic inputs follow the reference recipe (utils/src/embedding_allocation.cu
:113-168): table U(-1,1), power-law indices unique per bag, CSR bag lengths
U{0..hotness}, weights in {0.5, 0.25}, integer grad_y in [-10, 10].
`--check_result`: the CPU checker is test infrastructure (oracle/) and is not
reachable from here; tests/test_manual_benchmark.py runs this same pipeline
through `run(flags, check=...)` and compares every stage with the oracle.
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FLAGS = {  # name: (type, default)  -- benchmarks/manual_benchmark.cu:44-81
    "num_categories": (int, 1048576), "embed_width": (int, 128), "batch_size": (int, 1024),
    "hotness": (int, 1), "iterations": (int, 1), "alpha": (float, 0.0),
    "use_int64_indices": (bool, False), "check_result": (bool, False),
    "half_embedding_type": (bool, False), "csr_input": (bool, False),
    "weighted_sum": (bool, False), "fp16_math": (bool, False),
    "compressed_grad": (bool, True), "skip_grad_init": (bool, True),
    "forward_only": (bool, False), "enable_csv": (bool, False),
    "enable_stderr": (bool, True), "clear_caches": (bool, True),
    # extras
    "bf16": (bool, False), "combine_mode": (str, "sum"),
    "csv_file": (str, "manual_benchmark_out.csv"),
}
CSV_HEADER = ("num_categories,batch_size,hotness,alpha,embed_width,combine_mode,"
              "is_csr,is_weighted,compressed_grad,skip_grad_init,name,"
              "iterations,elapsed_time_ms,avg_time_ms,algo_bw_l2,algo_bw_dram")


def parse_flags(argv):
    """absl::ParseCommandLine for the flag set above."""
    vals = {k: d for k, (_, d) in FLAGS.items()}
    i = 0
    while i < len(argv):
        a = argv[i]
        if not a.startswith("-"):
            raise SystemExit(f"unexpected argument {a!r}")
        a = a.lstrip("-")
        name, eq, v = a.partition("=")
        if name.startswith("no") and name[2:] in FLAGS and FLAGS[name[2:]][0] is bool and not eq:
            vals[name[2:]] = False
            i += 1
            continue
        if name not in FLAGS:
            raise SystemExit(f"Unknown command line flag '{name}'")
        ty = FLAGS[name][0]
        if ty is bool:
            if eq:
                vals[name] = v.lower() in ("1", "true", "t", "yes", "y")
            elif i + 1 < len(argv) and argv[i + 1].lower() in ("true", "false"):
                vals[name] = argv[i + 1].lower() == "true"
                i += 1
            else:
                vals[name] = True
        else:
            if not eq:
                i += 1
                if i >= len(argv):
                    raise SystemExit(f"Missing the value for the flag '{name}'")
                v = argv[i]
            vals[name] = ty(v)
        i += 1
    return vals


def log(f, msg):
    if f["enable_stderr"]:
        print(msg, file=sys.stderr)


def combine_mode_str(mode):
    return {"sum": "kSum", "mean": "kMean", "concat": "kConcat"}[mode]


def csv_line(f, name, iterations, elapsed_ms, bw_l2, bw_dram):
    # benchmarks/manual_benchmark.cu:111-129 (including its trailing blanks)
    return (f"{f['num_categories']},{f['batch_size']},{f['hotness']},{f['alpha']:g},"
            f"{f['embed_width']},{combine_mode_str(f['combine_mode'])},{int(f['csr_input'])},"
            f"{int(f['weighted_sum'])},{int(f['compressed_grad'])},{int(f['skip_grad_init'])},"
            f"{name},{iterations} ,{elapsed_ms:.2f} ,{elapsed_ms / iterations:.2f} ,"
            f"{bw_l2:.2f},{bw_dram:.2f}")


def make_inputs(f, torch, dev):
    from benchmarks.sharded_bench import unique_bags_torch
    tdt = torch.float32
    if f["half_embedding_type"]:
        tdt = torch.float16
    if f["bf16"]:
        tdt = torch.bfloat16
    idt = torch.int64 if f["use_int64_indices"] else torch.int32
    rows, w, batch, hot = f["num_categories"], f["embed_width"], f["batch_size"], f["hotness"]
    g = torch.Generator(device=dev)
    g.manual_seed(123456)
    table = torch.empty(rows, w, dtype=tdt, device=dev)
    for r0 in range(0, rows, 1 << 20):
        r1 = min(rows, r0 + (1 << 20))
        table[r0:r1] = (torch.rand(r1 - r0, w, generator=g, device=dev) * 2 - 1).to(tdt)
    g.manual_seed(2024)
    bags = unique_bags_torch(g, batch, hot, rows, f["alpha"], dev)  # [batch, hot], distinct per bag
    offsets = None
    if f["csr_input"]:
        lens = torch.randint(0, hot + 1, (batch,), generator=g, device=dev)
        keep = torch.arange(hot, device=dev)[None, :] < lens[:, None]
        indices = bags[keep].to(idt).contiguous()
        offsets = torch.zeros(batch + 1, dtype=torch.int32, device=dev)
        offsets[1:] = torch.cumsum(lens, 0)
    else:
        indices = bags.reshape(-1).to(idt).contiguous()
    nnz = indices.numel()
    weights = None
    if f["weighted_sum"]:
        coin = torch.rand(nnz, generator=g, device=dev) < 0.5
        weights = torch.where(coin, 0.5, 0.25).to(tdt)
    g.manual_seed(654321)
    gy_rows = nnz if f["combine_mode"] == "concat" else batch
    grad_y = torch.randint(-10, 11, (gy_rows, w), generator=g, device=dev).to(tdt)
    return tdt, idt, table, indices, offsets, weights, grad_y, nnz


def main(argv=None):
    f = parse_flags(sys.argv[1:] if argv is None else argv)
    if f["check_result"]:
        raise SystemExit("--check_result: the CPU checker lives in tests/ "
                         "(pytest tests/test_manual_benchmark.py -m gpu)")
    return run(f)


def run(f, check=None):
    """Runs the benchmark.  `check(stage, tensors)` (tests only) is called after
    each stage with the device tensors that stage produced."""
    import torch
    import cuembed_b200 as ce
    if not torch.cuda.is_available():
        raise SystemExit("manual_benchmark needs a CUDA device: there is no CPU path")
    dev = torch.device("cuda", 0)
    log(f, "parsed flag " + ", ".join(f"{k}: {f[k]}" for k in FLAGS))
    mode = {"sum": ce.CombineMode.kSum, "mean": ce.CombineMode.kMean,
            "concat": ce.CombineMode.kConcat}[f["combine_mode"]]
    tdt, idt, table, indices, offsets, weights, grad_y, nnz = make_inputs(f, torch, dev)
    rows, w, batch, hot = f["num_categories"], f["embed_width"], f["batch_size"], f["hotness"]
    num_hots = 0 if f["csr_input"] else hot
    es, isz = table.element_size(), indices.element_size()
    iters = f["iterations"]
    out_rows = nnz if f["combine_mode"] == "concat" else batch
    out = torch.empty(out_rows, w, dtype=tdt, device=dev)
    flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev) if f["clear_caches"] else None
    outfile = None
    if f["enable_csv"]:
        new = not os.path.exists(f["csv_file"])
        outfile = open(f["csv_file"], "a")
        if new:
            outfile.write(CSV_HEADER + "\n")

    def timed(fn):
        """:199-248: warm-up, then per-iteration event pairs, caches cleared between."""
        fn()
        if flush is not None:
            flush.fill_(1)
        total = 0.0
        for _ in range(iters):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
            if flush is not None:
                flush.fill_(1)
        return total

    # ------------------------------------------------------------- forward
    def forward():
        ce.EmbeddingForward(table, w, indices, offsets, weights, batch, num_hots, mode, out,
                            fp16_math=f["fp16_math"])

    ms = timed(forward)
    if f["csr_input"]:
        bw = es * iters * (nnz - 1 + batch) * w / 1e6 / ms           # :250-254
    else:
        bw = es * iters * batch * (hot + (1 if f["combine_mode"] == "sum" else hot)) * w / 1e6 / ms
    if outfile:
        outfile.write(csv_line(f, "forward", iters, ms, bw, 0.0) + "\n")
    log(f, f"Embedding forward. Iterations: {iters} , Total time [ms]: {ms:.2f} , "
           f"Avg [ms]: {ms / iters:.2f} , Application BW [GB/s]: {bw:.2f}")
    if check is not None:
        check("forward", dict(table=table, indices=indices, offsets=offsets, weights=weights,
                              grad_y=grad_y, out=out, batch=batch, num_hots=num_hots,
                              mode=int(mode), width=w))
    if f["forward_only"]:
        return 0

    # ----------------------------------------------------------- transpose
    row_ids = torch.empty(nnz, dtype=idt, device=dev)
    t_idx = torch.empty(nnz, dtype=idt, device=dev)
    t_sid = torch.empty(nnz, dtype=idt, device=dev)
    t_w = torch.empty(nnz, dtype=tdt, device=dev) if weights is not None else None
    remapped = torch.empty(nnz, dtype=idt, device=dev) if f["compressed_grad"] else None
    lwork = max(ce.Transpose(row_ids, indices, weights, nnz, None, None, None, None),
                ce.ComputeCompressedGradIndices(indices, nnz, None, None))
    work = torch.empty(lwork, dtype=torch.uint8, device=dev)

    def transpose():  # utils/src/embedding_gpu_transpose.cu:32-79
        if f["combine_mode"] == "concat":
            ce.ExtractRowIdsForConcat(nnz, row_ids)
        elif f["csr_input"]:
            ce.ExtractRowIdsFromCSR(offsets, batch, row_ids)
        else:
            ce.ExtractRowIdsFromFixed(batch, hot, row_ids)
        ce.Transpose(row_ids, indices, weights, nnz, t_idx, t_sid, t_w, work)
        if remapped is not None:
            ce.ComputeCompressedGradIndices(t_idx, nnz, remapped, work)

    ms = timed(transpose)
    by = nnz * isz + (nnz * 4 if f["csr_input"] else 0) + (nnz * es if weights is not None else 0)
    by += (3 if f["compressed_grad"] else 2) * nnz * isz + (nnz * es if weights is not None else 0)
    bw = by * iters / 1e6 / ms                                           # :340-354
    if outfile:
        outfile.write(csv_line(f, "transpose", iters, ms, bw, 0.0) + "\n")
    log(f, f"Transpose. Iterations: {iters} , Total time [ms]: {ms:.2f} , "
           f"Avg [ms]: {ms / iters:.2f} , Application BW [GB/s]: {bw:.2f}")
    if check is not None:
        check("transpose", dict(t_idx=t_idx, t_sid=t_sid, t_w=t_w, remapped=remapped))

    # ------------------------------------------------------------ backward
    num_unique = int(torch.unique_consecutive(t_idx).numel())           # :447-449
    grad_rows = num_unique if f["compressed_grad"] else rows
    grad = torch.zeros(grad_rows, w, dtype=tdt, device=dev)
    inv = torch.empty(num_unique, dtype=idt, device=dev) if f["compressed_grad"] else None
    bwork = torch.empty(ce.backward_workspace_bytes(tdt, w, nnz, idt), dtype=torch.uint8,
                        device=dev)

    def backward():
        ce.EmbeddingBackward(grad_y, w, grad_rows, nnz, t_idx, t_sid, remapped, t_w,
                             f["skip_grad_init"], grad, inv, work=bwork)

    ms = timed(backward)
    dram = es * w * num_unique + isz * nnz * 2 + (es * nnz if weights is not None else 0)
    if f["combine_mode"] == "concat":
        dram += es * w * nnz
        l2 = dram
    else:
        dram += es * w * batch
        l2 = dram + es * w * nnz
    bw_dram, bw_l2 = dram * iters / 1e6 / ms, l2 * iters / 1e6 / ms    # :444-473
    if outfile:
        outfile.write(csv_line(f, "backward", iters, ms, bw_l2, bw_dram) + "\n")
    log(f, f"Backward. Iterations: {iters} , Total time [ms]: {ms:.2f} , "
           f"Avg [ms]: {ms / iters:.2f} , Application BW L2 [GB/s]: {bw_l2:.2f}"
           f", Application BW DRAM [GB/s]: {bw_dram:.2f}")
    if check is not None:
        check("backward", dict(grad=grad, inv=inv, grad_rows=grad_rows))
    if outfile:
        outfile.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
