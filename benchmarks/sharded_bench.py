"""bench.py --gpus N (N > 1): the row-sharded path, one process per GPU.
(Bench harness: lives next to the other benchmark drivers, not in the package.)

Weak scaling of the headline workload: every rank owns a C2-sized shard
(10 M x 256 fp16 rows), the global table has N x 10 M rows, the global batch
N x 65536 bags of hotness 64 drawn from the power law over the GLOBAL table
(replicated on every rank).  One step =
    forward : shard select -> local pool (fp32 partial) -> NCCL reduce-scatter
              -> epilogue
    transpose: row ids + sort + compressed remap of the rank's own lookups
    backward: NCCL all-gather of grad_y -> local backward (compressed)
Stage times are CUDA-event times, max over ranks, L2 flushed before every
stage.  value = global lookups / step time.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def power_law_torch(gen, n, num_categories, alpha, device):
    u = torch.rand(n, generator=gen, device=device, dtype=torch.float64)
    if alpha == 0.0:
        y = u * num_categories + 1.0
    else:
        g = 1.0 - alpha
        hi = float(num_categories + 1) ** g
        y = (u * (hi - 1.0) + 1.0) ** (1.0 / g)
    return y.floor().clamp_(1, num_categories).to(torch.int64)


def unique_bags_torch(gen, batch, hot, num_categories, alpha, device):
    """GPU version of cuembed_b200.datagen.unique_bags (same distribution)."""
    n_cat = num_categories - 1
    bags = power_law_torch(gen, batch * hot, n_cat, alpha, device).view(batch, hot)
    while True:
        bags, _ = torch.sort(bags, dim=1)
        dup = torch.zeros_like(bags, dtype=torch.bool)
        dup[:, 1:] = bags[:, 1:] == bags[:, :-1]
        n_dup = int(dup.sum().item())
        if n_dup == 0:
            break
        bags[dup] = power_law_torch(gen, n_dup, n_cat, alpha, device)
    perm = torch.randperm(n_cat + 1, generator=gen, device=device)
    bags = perm[bags]
    keys = torch.rand(bags.shape, generator=gen, device=device)
    order = torch.argsort(keys, dim=1)
    return torch.gather(bags, 1, order)


def run(args, rank, local_rank, world):
    """Dispatch on --transport: "p2p" (default: exchange fused into the kernels
    over peer memory) or "nccl" (reduce-scatter / all-gather baseline).  If the
    ranks cannot map each other's memory the p2p request falls back to nccl and
    says so in the JSON line."""
    transport = getattr(args, "transport", "p2p")
    note = None
    if transport == "p2p":
        ok = torch.tensor([1], device=torch.device("cuda", local_rank))
        try:
            from cuembed_b200 import peer
            probe = peer.PeerBuffer(4096)
            probe.close()
        except Exception as e:  # noqa: BLE001 -- any failure means "cannot map"
            ok.zero_()
            note = f"{type(e).__name__}: {e}"
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            return run_p2p(args, rank, local_rank, world)
        note = f"nccl (peer mapping unavailable: {note})"
    return run_nccl(args, rank, local_rank, world, note)


def make_inputs(args, rank, world, dev):
    import bench
    from cuembed_b200.sharded import row_range
    cfg = dict(bench.WORKLOADS[args.workload])
    tdt = {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32}[cfg["dtype"]]
    idt = torch.int32 if cfg["index"] == "int32" else torch.int64
    shard_rows, w, hot = cfg["num_categories"], cfg["embed_width"], cfg["hotness"]
    if cfg.get("global_problem"):
        rows, batch = cfg["num_categories"], cfg["batch_size"]      # strong scaling
    else:
        rows, batch = shard_rows * world, cfg["batch_size"] * world  # weak scaling
    nnz = batch * hot
    lo, hi = row_range(rows, world, rank)
    indices = torch.empty(nnz, dtype=idt, device=dev)
    if rank == 0:
        g = torch.Generator(device=dev)
        g.manual_seed(1234)
        indices.copy_(unique_bags_torch(g, batch, hot, rows, cfg["alpha"], dev).view(-1).to(idt))
    dist.broadcast(indices, src=0)
    g = torch.Generator(device=dev)
    g.manual_seed(123456 + rank)
    table = torch.empty(hi - lo, w, dtype=tdt, device=dev)
    for r0 in range(0, hi - lo, 1 << 24):  # U(-1, 1) in place, 16 Mi rows at a time
        table[r0:min(hi - lo, r0 + (1 << 24))].uniform_(-1.0, 1.0, generator=g)
    per = batch // world
    g.manual_seed(654321 + rank)
    grad_slice = torch.randint(-10, 11, (per, w), generator=g, device=dev).to(tdt)
    return cfg, tdt, idt, rows, batch, nnz, per, indices, table, grad_slice


def verify_p2p(emb, ce, table, indices, grad_slice, batch, hot, world, rank, dev):
    """Real-rank numerics check, run AFTER the timed region: the sharded
    forward / backward against an independent computation made of plain torch
    ops + NCCL collectives.  The table is replaced in place by round(8 * table)
    (integers in [-8, 8]) and grad_y is the integer recipe, so every partial and
    every sum is exact in fp32 AND in 16-bit partials: the comparison is
    bit-exact whatever the summation order or the wire type.
      forward : the first 512 bags of every rank's slice; each rank adds the
                rows it owns, NCCL all_reduce sums the ranks
      backward: the whole compressed gradient of this rank's shard against
                index_add over an NCCL all_gather of grad_y, and the row list"""
    w = table.shape[1]
    lo, hi = emb.lo, emb.hi
    per = batch // world
    nnz = batch * hot
    table.mul_(8).round_()
    out, ctx = emb.forward(indices, None, None, batch, hot, ce.CombineMode.kSum)
    nb = min(per, 512)
    sample = torch.cat([torch.arange(r * per, r * per + nb, device=dev) for r in range(world)])
    idx = indices.view(batch, hot)[sample].long()
    own = (idx >= lo) & (idx < hi)
    local = (idx - lo).clamp_(0, max(hi - lo - 1, 0))
    want = torch.zeros(world * nb, w, dtype=torch.float32, device=dev)
    for j in range(hot):  # one lookup column at a time keeps the temporaries small
        want += table[local[:, j]].float() * own[:, j, None]
    dist.all_reduce(want)
    ok_f = bool(torch.equal(out[:nb].float(), want[rank * nb:(rank + 1) * nb]))
    grad, rows_out = emb.backward(grad_slice, ctx, True)
    full_gy = torch.empty(batch, w, dtype=grad_slice.dtype, device=dev)
    dist.all_gather_into_tensor(full_gy, grad_slice.contiguous())
    flat = indices.long()
    mine = (flat >= lo) & (flat < hi)
    li = flat[mine]
    lb = (torch.arange(nnz, device=dev) // hot)[mine]
    uniq, inverse = torch.unique(li, return_inverse=True)
    want_g = torch.zeros(uniq.numel(), w, dtype=torch.float32, device=dev)
    for n0 in range(0, li.numel(), 1 << 19):
        n1 = min(li.numel(), n0 + (1 << 19))
        want_g.index_add_(0, inverse[n0:n1], full_gy[lb[n0:n1]].float())
    ok_b = (grad.shape[0] == uniq.numel()
            and bool(torch.equal(grad, want_g.to(grad.dtype)))
            and bool(torch.equal(rows_out.long(), uniq)))
    torch.cuda.synchronize()
    ok = torch.tensor([int(ok_f), int(ok_b), int(emb.status() == 0)], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    f, b, st = (bool(x) for x in ok.tolist())
    return {"verified": f and b and st, "forward": f, "backward": b, "no_wait_timeout": st,
            "how": "after the timed region, every rank: integer-valued table (round(8 * table)) "
                   "and integer grad_y -> exact sums; forward of 512 bags per rank slice vs "
                   "torch gather + NCCL all_reduce, whole compressed shard gradient + row list vs "
                   "index_add over an NCCL all_gather; bit-exact; AND over ranks"}


def run_p2p(args, rank, local_rank, world):
    """One step = forward (pool + push over NVLink, signal, rank-ordered reduce)
    -> [copy engines push grad_y slices || select + sort of the own lookups]
    -> wait -> backward.  The WHOLE step is timed with one CUDA-event pair (L2
    flushed and ranks aligned before it), max over ranks; the split into stages
    comes from events recorded inside the step."""
    import bench
    import cuembed_b200 as ce
    from cuembed_b200.sharded_p2p import PeerShardedEmbedding

    dev = torch.device("cuda", local_rank)
    cfg, tdt, idt, rows, batch, nnz, per, indices, table, grad_slice = \
        make_inputs(args, rank, world, dev)
    w, hot = cfg["embed_width"], cfg["hotness"]
    shard_rows = cfg["num_categories"]
    pdt = {"f32": torch.float32, "table": tdt}[getattr(args, "partial_dtype", "table")]
    sf = {"auto": None, "on": True, "off": False}[getattr(args, "select_first", "auto")]
    emb = PeerShardedEmbedding(table, rows, partial_dtype=pdt, select_first=sf)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    # sizes that a training loop knows from the previous identical step
    out, ctx = emb.forward(indices, None, None, batch, hot, ce.CombineMode.kSum)
    emb.prepare_backward(ctx, True)
    local_nnz = ctx.local_nnz
    num_unique = int(ctx.coo[3][-1].item()) + 1
    grad = torch.zeros(num_unique, w, dtype=tdt, device=dev)
    inv = torch.empty(num_unique, dtype=idt, device=dev)
    g2, _ = emb.backward(grad_slice, ctx)
    torch.cuda.synchronize()
    state = {}
    known_sizes = bool(getattr(args, "known_sizes", False))

    def one_step(rec=None):
        flush.fill_(1)
        dist.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(stream)
        out, ctx = emb.forward(indices, None, None, batch, hot, ce.CombineMode.kSum,
                               local_nnz=local_nnz if known_sizes else None)
        ev[1].record(stream)
        if known_sizes:
            pend = emb.backward_begin(grad_slice, ctx, True, local_nnz=local_nnz)
            ev[2].record(stream)
            emb.backward_finish(pend, grad=grad, inverse=inv)
        else:
            # a loop whose indices change every step: the host reads the number
            # of owned lookups (after the select) and the number of unique rows
            # (after the sort) back in EVERY step to size the sort / gradient
            pend = emb.backward_begin(grad_slice, ctx, True)
            ev[2].record(stream)
            emb.backward_finish(pend)
        ev[3].record(stream)
        state["out"] = out
        if rec is not None:
            rec.append(ev)

    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()
    dist.barrier()
    if getattr(args, "trace", None):
        # CUPTI timeline of a few steps (not a bench number): per-kernel GPU time
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                one_step()
            torch.cuda.synchronize()
        if rank == 0:
            with open(args.trace, "w") as f:
                f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60,
                                                  max_name_column_width=90))
        dist.barrier()
    sampler = bench.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ce.launch_count()
    events = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step(events)
    torch.cuda.synchronize()
    dist.barrier()
    wall = time.perf_counter() - t0
    launches = ce.launch_count() - launches0
    acc = [0.0, 0.0, 0.0, 0.0]
    for ev in events:
        acc[0] += ev[0].elapsed_time(ev[3])
        acc[1] += ev[0].elapsed_time(ev[1])
        acc[2] += ev[1].elapsed_time(ev[2])
        acc[3] += ev[2].elapsed_time(ev[3])
    t = torch.tensor([a / args.steps for a in acc], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step, fwd_ms, tr_ms, bwd_ms = (float(x) for x in t.tolist())
    status = emb.status()

    # end to end with host buffers: indices + grad slice in, output slice +
    # compressed gradient + row list out
    e2e_ms = float("nan")
    h2d = d2h = 0
    if not args.no_e2e:
        idx_host = indices.cpu().pin_memory()
        gy_host = grad_slice.cpu().pin_memory()
        out_host = torch.empty(per, w, dtype=tdt).pin_memory()
        grad_host = torch.empty(num_unique, w, dtype=tdt).pin_memory()
        inv_host = torch.empty(num_unique, dtype=idt).pin_memory()

        def e2e_step():
            indices.copy_(idx_host, non_blocking=True)
            out, ctx = emb.forward(indices, None, None, batch, hot, ce.CombineMode.kSum)
            out_host.copy_(out, non_blocking=True)
            grad_slice.copy_(gy_host, non_blocking=True)
            pend = emb.backward_begin(grad_slice, ctx, True, local_nnz=local_nnz)
            emb.backward_finish(pend, grad=grad, inverse=inv)
            grad_host.copy_(grad, non_blocking=True)
            inv_host.copy_(inv, non_blocking=True)

        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        torch.cuda.synchronize()
        c = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.MAX)
        e2e_ms = float(c.item())
        es = table.element_size()
        h2d = world * (nnz * indices.element_size() + per * w * es)
        d2h = world * (per * w * es + num_unique * (w * es + indices.element_size()))

    clocks = sampler.stop() if rank == 0 else None
    verified = verify_p2p(emb, ce, table, indices, grad_slice, batch, hot, world, rank, dev)
    peak, peak_src = bench.measured_peaks()
    es = table.element_size()
    isz = indices.element_size()
    psz = torch.empty(0, dtype=pdt).element_size()
    # per-rank algorithmic bytes of the local kernels (reference accounting):
    # rows gathered + partial rows written / received + gradient traffic
    fwd_bytes = es * w * local_nnz + psz * w * batch + psz * w * per * world + es * w * per
    bwd_bytes = es * w * (local_nnz + batch + num_unique) + 2 * isz * local_nnz
    nvlink_fwd = psz * w * per * (world - 1)
    nvlink_bwd = es * w * per * (world - 1)
    if rank == 0:
        achieved = (fwd_bytes + bwd_bytes) / ((fwd_ms + bwd_ms) * 1e-3) / 1e9
        line = {
            "metric": "lookups/s (fwd + transpose + bwd, compressed grad)",
            "value": round(nnz / (ms_per_step * 1e-3), 1), "unit": "lookups/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "verified": verified["verified"], "verification": verified,
            "scaling": "strong" if cfg.get("global_problem") else "weak",
            "vs_baseline": None, "dtype": cfg["dtype"],
            "data": "synthetic",
            "config": {"workload": f"row-sharded {args.workload}: {world} shards of "
                                   f"{table.shape[0]}x{w} {cfg['dtype']} (global {rows} rows), global batch "
                                   f"{batch}, hotness {hot}, alpha {cfg['alpha']}, {cfg['index']} indices, "
                                   f"sum, compressed grad",
                       "parallelism": f"row-sharded x{world}, exchange fused over NVLink peer memory: "
                                      f"pool kernel stores {('fp32' if psz == 4 else cfg['dtype'])} partial "
                                      f"rows into the bag owner's slots, rank-ordered reduce; grad_y "
                                      f"slices pushed by copy engines during the local sort",
                       "transport": "p2p",
                       "l2": "flushed before every step (512 MB write)",
                       "select_first": bool(emb.select_first),
                       "sizes": ("local nnz / num_unique known from an earlier identical step "
                                 "(--known-sizes): no host read-back" if known_sizes else
                                 "local nnz and num_unique read back by the host in every step, "
                                 "inside the timed region: local nnz is summed from the pool "
                                 "kernel's counts and copied on the side stream under the exchange; "
                                 "num_unique is read after the backward has been launched into "
                                 "upper-bound buffers"),
                       "nnz_global": nnz, "nnz_local_rank0": local_nnz,
                       "num_unique_rank0": num_unique,
                       "nvlink_bytes_out_per_rank": {"forward": nvlink_fwd, "backward": nvlink_bwd},
                       "peer_wait_status": status},
            "stages": {"forward": {"ms": round(fwd_ms, 4)},
                       ("sort_with_grad_push" if emb.select_first else
                        "select_sort_with_grad_push"): {"ms": round(tr_ms, 4)},
                       "backward": {"ms": round(bwd_ms, 4)}},
            "roofline": {"bound": "hbm", "kernel": "ShardPoolPushKernel + BwdWarpKernel (local)",
                         "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": None,
                         "peak_source": peak_src,
                         "note": "per-rank algorithmic bytes of forward+backward over their in-step "
                                 "time incl. the NVLink exchange"},
            "cpu_baseline": None,
            "e2e": {"value": (round(nnz / (e2e_ms * 1e-3), 1) if e2e_ms == e2e_ms else None),
                    "unit": "lookups/s",
                    "ms_per_step": (round(e2e_ms, 4) if e2e_ms == e2e_ms else None),
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches) * world,
            "clocks": clocks,
            "wall_s_timed_region": round(wall, 3),
        }
        print(json.dumps(line))
    dist.barrier()
    emb.close()
    dist.destroy_process_group()


def run_nccl(args, rank, local_rank, world, transport_note=None):
    import bench
    import cuembed_b200 as ce
    from cuembed_b200.sharded import RowShardedEmbedding, row_range

    dev = torch.device("cuda", local_rank)
    cfg = dict(bench.WORKLOADS[args.workload])
    tdt = {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32}[cfg["dtype"]]
    idt = torch.int32 if cfg["index"] == "int32" else torch.int64
    shard_rows, w, hot = cfg["num_categories"], cfg["embed_width"], cfg["hotness"]
    rows = shard_rows * world
    batch = cfg["batch_size"] * world
    nnz = batch * hot
    lo, hi = row_range(rows, world, rank)

    # replicated indices: rank 0 generates, everyone receives
    indices = torch.empty(nnz, dtype=idt, device=dev)
    if rank == 0:
        g = torch.Generator(device=dev)
        g.manual_seed(1234)
        indices.copy_(unique_bags_torch(g, batch, hot, rows, cfg["alpha"], dev).view(-1).to(idt))
    dist.broadcast(indices, src=0)
    g = torch.Generator(device=dev)
    g.manual_seed(123456 + rank)
    table = torch.empty(hi - lo, w, dtype=tdt, device=dev)
    for r0 in range(0, hi - lo, 1 << 20):
        r1 = min(hi - lo, r0 + (1 << 20))
        table[r0:r1] = (torch.rand(r1 - r0, w, generator=g, device=dev) * 2 - 1).to(tdt)
    per = batch // world
    g.manual_seed(654321 + rank)
    grad_slice = torch.randint(-10, 11, (per, w), generator=g, device=dev).to(tdt)

    emb = RowShardedEmbedding(table, rows)
    ops = emb.ops
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    state = {}

    def forward():
        out, ctx = emb.forward(indices, None, None, batch, hot, ce.CombineMode.kSum)
        state["out"], state["ctx"] = out, ctx

    forward()
    ctx = state["ctx"]
    local_nnz = int(ctx.local_offsets[-1].item())
    row_ids = torch.empty(local_nnz, dtype=idt, device=dev)
    t_idx = torch.empty(local_nnz, dtype=idt, device=dev)
    t_sid = torch.empty(local_nnz, dtype=idt, device=dev)
    remapped = torch.empty(local_nnz, dtype=idt, device=dev)
    l_idx = ctx.local_indices[:local_nnz]
    lwork = max(ce.Transpose(row_ids, l_idx, None, local_nnz, None, None, None, None),
                ce.ComputeCompressedGradIndices(l_idx, local_nnz, None, None))
    work = torch.empty(lwork, dtype=torch.uint8, device=dev)

    def transpose():
        ce.ExtractRowIdsFromCSR(state["ctx"].local_offsets, batch, row_ids)
        ce.Transpose(row_ids, l_idx, None, local_nnz, t_idx, t_sid, None, work)
        ce.ComputeCompressedGradIndices(t_idx, local_nnz, remapped, work)

    transpose()
    num_unique = int(remapped[-1].item()) + 1
    grad = torch.zeros(num_unique, w, dtype=tdt, device=dev)
    inv = torch.empty(num_unique, dtype=idt, device=dev)
    bwork = torch.empty(ce.backward_workspace_bytes(tdt, w, local_nnz, idt),
                        dtype=torch.uint8, device=dev)
    full_gy = torch.empty(batch, w, dtype=tdt, device=dev)

    def backward():
        dist.all_gather_into_tensor(full_gy, grad_slice)
        ce.EmbeddingBackward(full_gy, w, num_unique, local_nnz, t_idx, t_sid, remapped,
                             None, True, grad, inv, work=bwork)

    stages = [("forward", forward), ("transpose", transpose), ("backward", backward)]
    stream = torch.cuda.current_stream()

    def one_step(times=None):
        for name, fn in stages:
            flush.fill_(1)
            dist.barrier()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            if times is not None:
                times.append((name, e0, e1))

    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = bench.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ce.launch_count()
    events = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step(events)
    torch.cuda.synchronize()
    dist.barrier()
    wall = time.perf_counter() - t0
    launches = ce.launch_count() - launches0
    per_stage = {"forward": 0.0, "transpose": 0.0, "backward": 0.0}
    for name, e0, e1 in events:
        per_stage[name] += e0.elapsed_time(e1)
    t = torch.tensor([per_stage[k] / args.steps for k in ("forward", "transpose", "backward")],
                     dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks, per stage
    fwd_ms, tr_ms, bwd_ms = (float(x) for x in t.tolist())
    ms_per_step = fwd_ms + tr_ms + bwd_ms

    # communication alone (same buffers), for the report
    partial = torch.empty(batch, w, dtype=torch.float32, device=dev)
    mine = torch.empty(per, w, dtype=torch.float32, device=dev)
    comm = {}
    for name, fn in (("reduce_scatter", lambda: dist.reduce_scatter_tensor(mine, partial)),
                     ("all_gather", lambda: dist.all_gather_into_tensor(full_gy, grad_slice))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        c = torch.tensor([e0.elapsed_time(e1) / 10], dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.MAX)
        comm[name] = float(c.item())

    # end to end with host buffers: indices in, output slice + gradient out
    e2e_ms = float("nan")
    h2d = d2h = 0
    if not args.no_e2e:
        idx_host = indices.cpu().pin_memory()
        gy_host = grad_slice.cpu().pin_memory()
        out_host = torch.empty(per, w, dtype=tdt).pin_memory()
        grad_host = torch.empty(num_unique, w, dtype=tdt).pin_memory()
        inv_host = torch.empty(num_unique, dtype=idt).pin_memory()

        def e2e_step():
            indices.copy_(idx_host, non_blocking=True)
            forward()
            out_host.copy_(state["out"], non_blocking=True)
            grad_slice.copy_(gy_host, non_blocking=True)
            transpose()
            backward()
            grad_host.copy_(grad, non_blocking=True)
            inv_host.copy_(inv, non_blocking=True)

        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        torch.cuda.synchronize()
        c = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.MAX)
        e2e_ms = float(c.item())
        es = table.element_size()
        h2d = world * (nnz * indices.element_size() + per * w * es)
        d2h = world * (per * w * es + num_unique * (w * es + indices.element_size()))

    clocks = sampler.stop() if rank == 0 else None
    peak, peak_src = bench.measured_peaks()
    es = table.element_size()
    isz = indices.element_size()
    # per-rank algorithmic bytes of the local kernels (reference accounting)
    fwd_bytes = es * w * local_nnz + 4 * w * batch      # rows gathered + fp32 partial written
    bwd_bytes = es * w * (local_nnz + batch + num_unique) + 2 * isz * local_nnz
    if rank == 0:
        line = {
            "metric": "lookups/s (fwd + transpose + bwd, compressed grad)",
            "value": round(nnz / (ms_per_step * 1e-3), 1), "unit": "lookups/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": cfg["dtype"],
            "data": "synthetic",
            "config": {"workload": f"row-sharded manual_benchmark shape: {world} shards of "
                                   f"{shard_rows}x{w} {cfg['dtype']} (global {rows} rows), global batch "
                                   f"{batch}, hotness {hot}, alpha {cfg['alpha']}, {cfg['index']} indices, "
                                   f"sum, compressed grad",
                       "parallelism": f"row-sharded x{world}: NCCL reduce-scatter of fp32 partial sums "
                                      f"(forward), all-gather of grad_y (backward)",
                       "transport": transport_note or "nccl",
                       "l2": "flushed before every stage (512 MB write)",
                       "nnz_global": nnz, "nnz_local_rank0": local_nnz, "num_unique_rank0": num_unique},
            "stages": {"forward": {"ms": round(fwd_ms, 4)}, "transpose": {"ms": round(tr_ms, 4)},
                       "backward": {"ms": round(bwd_ms, 4)},
                       "collectives_alone_ms": {k: round(v, 4) for k, v in comm.items()}},
            "roofline": {"bound": "hbm", "kernel": "BwdWarpKernel (local) + collectives",
                         "achieved": round((fwd_bytes + bwd_bytes) / ((fwd_ms + bwd_ms) * 1e-3) / 1e9, 1),
                         "peak": peak, "unit": "GB/s",
                         "frac": round((fwd_bytes + bwd_bytes) / ((fwd_ms + bwd_ms) * 1e-3) / 1e9 / peak, 4),
                         "traffic": None, "peak_source": peak_src,
                         "note": "per-rank algorithmic bytes of forward+backward over their stage time "
                                 "incl. the collectives"},
            "cpu_baseline": None,
            "e2e": {"value": (round(nnz / (e2e_ms * 1e-3), 1) if e2e_ms == e2e_ms else None),
                    "unit": "lookups/s",
                    "ms_per_step": (round(e2e_ms, 4) if e2e_ms == e2e_ms else None),
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches) * world,
            "clocks": clocks,
            "wall_s_timed_region": round(wall, 3),
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
