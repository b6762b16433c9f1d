"""Row-sharded mode with the exchange fused over peer memory
(cuembed_b200/sharded_p2p.py, csrc/sharded_p2p.cu).

* `not gpu`: the ABI carries the peer entry points; constants agree.
* `gpu`, one device: several VIRTUAL ranks in one process (peer.LocalPeerGroup):
  the kernels store into each other's buffers exactly as they would over NVLink,
  so pool-push, signal / wait, rank-ordered reduce, concat push, all-gather push
  and the epoch double buffering are checked against the CPU oracle.
* `gpu`, two devices (skipped on a one-GPU box): two processes, CUDA-IPC mapped
  buffers, real peer stores.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
from cuembed_b200.api import CombineMode
from cuembed_b200.sharded import row_range
from helpers import Problem, to_f32
from oracle.cpu_lib import BF16, F16, F32


def test_peer_abi_symbols_and_constants():
    import ctypes
    import re
    from cuembed_b200 import _lib, build, peer
    lib = ctypes.CDLL(build.build())
    for name in ("cuembed_peer_alloc", "cuembed_peer_free", "cuembed_peer_export",
                 "cuembed_peer_open", "cuembed_peer_close", "cuembed_shard_pool_push",
                 "cuembed_shard_concat_push", "cuembed_shard_signal", "cuembed_shard_wait",
                 "cuembed_shard_reduce_finalize", "cuembed_shard_allgather_push"):
        assert hasattr(lib, name), name
        assert name in _lib.SIGNATURES
    hdr = open(os.path.join(os.path.dirname(build.__file__), "..", "include",
                            "cuembed_b200.h")).read()
    consts = dict(re.findall(r"#define (CUEMBED_(?:MAX_WORLD|PEER_\w+)) (\d+)", hdr))
    assert int(consts["CUEMBED_MAX_WORLD"]) == peer.MAX_WORLD
    assert int(consts["CUEMBED_PEER_CHANNELS"]) == peer.CHANNELS
    assert int(consts["CUEMBED_PEER_FLAG_BYTES"]) == peer.FLAG_BYTES
    assert int(consts["CUEMBED_PEER_HANDLE_BYTES"]) == peer.HANDLE_BYTES
    assert peer.FLAG_BYTES >= 4 * (peer.CHANNELS * peer.MAX_WORLD + 1)


# ------------------------------------------------------------------ expected
def expected_forward(p, world, oracle, out_np_dtype=None):
    """Per-rank fp32 partial sums (sequential in bag order, the oracle's loop)
    added in RANK ORDER, then the epilogue: what the fused path must produce
    bit for bit."""
    from sharded_helpers import OracleLocalOps
    ops = OracleLocalOps()
    idx = torch.from_numpy(p.indices)
    off = torch.from_numpy(p.offsets) if p.offsets is not None else None
    w = torch.from_numpy(helpers.raw(p.weights)) if p.weights is not None else None
    total = None
    for r in range(world):
        lo, hi = row_range(p.num_categories, world, r)
        l_off, l_idx, l_w = ops.shard_select(idx, off, w, p.batch, p.num_hots, lo, hi)
        table = helpers.raw(p.table)[lo:hi]
        lw = l_w.numpy() if l_w is not None else None
        if p.dt == BF16:
            from oracle.cpu_lib import Bf16
            table = Bf16(np.ascontiguousarray(table))
            lw = Bf16(np.ascontiguousarray(lw)) if lw is not None else None
        if hi > lo:
            part = oracle.forward(table, l_idx.numpy(), l_off.numpy(), lw, p.batch, 0,
                                  helpers.SUM, embed_width=p.width, out_dt=F32)
        else:
            part = np.zeros((p.batch, p.width), np.float32)
        total = part if total is None else (total + part).astype(np.float32)
    if p.mode == helpers.MEAN:
        off_np = p.offsets if p.offsets is not None else np.arange(p.batch + 1) * p.num_hots
        for s in range(p.batch):
            a, b = int(off_np[s]), int(off_np[s + 1])
            if p.weights is not None:
                denom = np.float32(0)
                for j in range(a, b):
                    denom = np.float32(denom + to_f32(p.weights)[j])
            else:
                denom = np.float32(b - a)
            total[s] = 0 if denom == 0 else total[s] * np.float32(np.float32(1.0) / denom)
    return total


def run_virtual(p, world, oracle, partial_dtype=torch.float32, steps=1, compressed=True,
                select_first=None):
    """`steps` forward + backward rounds on `world` virtual ranks of cuda:0.
    Returns per-step lists of (outs, grads, rows)."""
    import gpu_helpers as gh
    from cuembed_b200 import peer
    from cuembed_b200.sharded_p2p import PeerShardedEmbedding
    dev = torch.device(gh.DEV)
    group = peer.LocalPeerGroup(world, dev)
    shared = {}

    def factory(rank):
        def make(kind, nbytes):
            key = (kind, nbytes)
            if key not in shared:
                shared[key] = group.alloc(nbytes)
            return shared[key][rank]
        return make

    table = gh.to_dev(p.table)
    embs = []
    for r in range(world):
        lo, hi = row_range(p.num_categories, world, r)
        embs.append(PeerShardedEmbedding(table[lo:hi].contiguous(), p.num_categories,
                                         partial_dtype=partial_dtype, rank=r, world=world,
                                         buffers=factory(r), select_first=select_first))
    idx, off, w = gh.to_dev(p.indices), gh.to_dev(p.offsets), gh.to_dev(p.weights)
    gy = gh.to_dev(p.grad_y)
    mode = CombineMode(p.mode)
    results = []
    try:
        for _ in range(steps):
            pend = [e.forward_begin(idx, off, w, p.batch, p.num_hots, mode) for e in embs]
            fwd = [e.forward_finish(x) for e, x in zip(embs, pend)]
            n = gy.shape[0] // world
            pend = [e.backward_begin(gy[r * n:(r + 1) * n], fwd[r][1], compressed)
                    for r, e in enumerate(embs)]
            bwd = [e.backward_finish(x) for e, x in zip(embs, pend)]
            torch.cuda.synchronize()
            assert all(e.status() == 0 for e in embs), "a peer wait timed out"
            results.append(([gh.to_host(o.clone()) for o, _ in fwd],
                            [gh.to_host(g) for g, _ in bwd],
                            [r.cpu().numpy() if r is not None else None for _, r in bwd],
                            [c.counts.cpu().numpy() if c.counts is not None else None
                             for _, c in fwd]))
    finally:
        torch.cuda.synchronize()
        for e in embs:
            e._bufs = {}
        group.close()
    return results


def check_backward(p, world, oracle, grads, rows):
    _, t_idx, t_sid, t_w, remapped = p.cpu_transpose(oracle)
    (g_all, inv_all), _ = p.cpu_backward(oracle, t_idx, t_sid, t_w, remapped, acc_f32=True)
    for r in range(world):
        lo, hi = row_range(p.num_categories, world, r)
        sel = (inv_all >= lo) & (inv_all < hi)
        assert np.array_equal(rows[r], inv_all[sel])
        assert helpers.value_equal(grads[r], helpers.raw(g_all)[sel] if p.dt != BF16
                                   else to_f32(g_all)[sel])


CASES = [
    # world, mode, csr, weighted, dt, index dtype, width, hot, batch
    (2, "sum", False, False, F16, np.int32, 256, 64, 64),
    (4, "sum", False, False, F32, np.int32, 32, 8, 96),
    (3, "mean", True, False, F32, np.int64, 24, 20, 99),
    (2, "sum", True, True, F32, np.int32, 16, 40, 130),
    (4, "mean", True, True, F16, np.int64, 128, 150, 64),
    (2, "mean", False, False, BF16, np.int32, 64, 7, 50),
    (8, "sum", False, False, F16, np.int32, 128, 64, 128),
    (2, "sum", False, False, F32, np.int32, 5, 3, 10),          # 4-byte vectors
    (2, "sum", True, False, F16, np.int32, 1024, 300, 16),      # column tiles, queue drains
]


@pytest.mark.gpu
@pytest.mark.parametrize("world,mode,csr,weighted,dt,it,width,hot,batch", CASES)
def test_virtual_ranks_forward_backward(cuda_lib, oracle, world, mode, csr, weighted, dt,
                                        it, width, hot, batch):
    p = Problem(batch, width, hot, mode, csr=csr, weighted=weighted, compressed=True,
                num_categories=997, dt=dt, index_dtype=it, alpha=1.05, seed=61)
    outs, grads, rows, counts = run_virtual(p, world, oracle)[0]
    want = expected_forward(p, world, oracle)
    got = np.concatenate([to_f32(o) for o in outs])
    want_cast = to_f32(helpers.cast_elems(want, dt))
    assert np.array_equal(got.view(np.uint32), want_cast.view(np.uint32))
    # within 1e-5 of the single-table oracle (association differs across shards)
    ref = to_f32(p.cpu_forward(oracle, out_dt=F32))
    tol = 1e-5 if dt == F32 else (2 ** -10 if dt == F16 else 2 ** -7)
    assert np.all(np.abs(got - ref) <= tol * np.maximum(1.0, np.abs(ref)) * 4)
    # counts = lookups per bag owned by the rank
    off = p.offsets if p.offsets is not None else np.arange(batch + 1) * hot
    for r in range(world):
        lo, hi = row_range(p.num_categories, world, r)
        keep = (p.indices >= lo) & (p.indices < hi)
        c = np.add.reduceat(np.concatenate([keep, [False]]).astype(np.int64), off[:-1])
        c[off[:-1] == off[1:]] = 0
        assert np.array_equal(counts[r], c)
    if dt != BF16:
        p2 = p
        check_backward(p2, world, oracle, grads, rows)


@pytest.mark.gpu
def test_virtual_ranks_repeated_steps_double_buffering(cuda_lib, oracle):
    p = Problem(63, 64, 16, "sum", compressed=True, num_categories=500, dt=F32, seed=63,
                integer_table=True)
    res = run_virtual(p, 3, oracle, steps=5)
    want = p.cpu_forward(oracle)
    for outs, grads, rows, _ in res:
        assert np.array_equal(np.concatenate(outs), want)
        check_backward(p, 3, oracle, grads, rows)


@pytest.mark.gpu
def test_virtual_ranks_16bit_partials(cuda_lib, oracle):
    p = Problem(64, 128, 32, "sum", num_categories=2000, dt=F16, alpha=1.15, seed=65,
                compressed=True)
    outs, _, _, _ = run_virtual(p, 4, oracle, partial_dtype=torch.float16)[0]
    got = np.concatenate([to_f32(o) for o in outs])
    ref = to_f32(p.cpu_forward(oracle, out_dt=F32))
    # every partial is rounded to fp16 once more: <= world/2 + 1/2 ulp of the terms
    assert np.all(np.abs(got - ref) <= 2 ** -10 * 4 * np.maximum(1.0, np.abs(ref)))


@pytest.mark.gpu
def test_virtual_ranks_select_first(cuda_lib, oracle):
    """select_first: the rank selects its own lookups before the forward and the
    pool-and-push kernel walks the selection (CSR bags of owned, rebased
    indices); identical results, forward and backward."""
    p = Problem(96, 64, 20, "mean", csr=True, weighted=True, compressed=True,
                num_categories=700, dt=F32, seed=69, integer_table=True)
    base = run_virtual(p, 4, oracle, select_first=False)[0]
    sel = run_virtual(p, 4, oracle, select_first=True, steps=2)
    for outs, grads, rows, _ in sel:
        for a, b in zip(outs, base[0]):
            assert helpers.bits_equal(a, b)
        for a, b in zip(grads, base[1]):
            assert helpers.bits_equal(a, b)
        for a, b in zip(rows, base[2]):
            assert np.array_equal(a, b)


@pytest.mark.gpu
def test_default_partial_type_is_the_table_type(cuda_lib, oracle):
    """fp16 / bf16 tables put 16-bit partial sums on the wire by default (half
    the NVLink bytes), fp32 tables fp32 partials.  Stated tolerance of the
    16-bit default against the fp32-accumulating oracle: every one of the
    `world` partials is rounded once to the 16-bit type before the rank-ordered
    fp32 reduction, i.e. |error| <= world * 2^-11 (fp16; 2^-8 for bf16) of the
    largest partial, plus the final rounding of the output."""
    for dt, eps in ((F16, 2.0 ** -11), (helpers.BF16, 2.0 ** -8)):
        p = Problem(64, 128, 32, "sum", num_categories=2000, dt=dt, alpha=1.15, seed=66,
                    compressed=True)
        world = 4
        outs, grads, rows, _ = run_virtual(p, world, oracle, partial_dtype=None)[0]
        got = np.concatenate([to_f32(o) for o in outs])
        ref = to_f32(p.cpu_forward(oracle, out_dt=F32))
        scale = np.maximum(1.0, np.abs(ref))
        assert np.all(np.abs(got - ref) <= (world + 1) * eps * 4 * scale)
        check_backward(p, world, oracle, grads, rows)  # gradients never touch the partials


@pytest.mark.gpu
def test_missing_peer_poisons_the_result_and_raises(cuda_lib):
    """A rank that never signals: the waiting kernel gives up after the peer
    timeout, fills the output with NaN instead of summing stale slots, and the
    next call on the object raises (silent stale sums)."""
    import gpu_helpers as gh
    from cuembed_b200 import peer
    from cuembed_b200.api import CuEmbedError
    from cuembed_b200.sharded_p2p import PeerShardedEmbedding
    dev = torch.device(gh.DEV)
    world = 2
    group = peer.LocalPeerGroup(world, dev)
    shared = {}

    def factory(rank):
        def make(kind, nbytes):
            key = (kind, nbytes)
            if key not in shared:
                shared[key] = group.alloc(nbytes)
            return shared[key][rank]
        return make

    table = torch.ones(100, 32, dtype=torch.float32, device=dev)
    embs = [PeerShardedEmbedding(table[r * 50:(r + 1) * 50].contiguous(), 100, rank=r,
                                 world=world, buffers=factory(r)) for r in range(world)]
    idx = torch.randint(0, 100, (8 * 4,), dtype=torch.int32, device=dev)
    try:
        embs[0].set_peer_timeout_ms(200)
        # only rank 0 runs the step: rank 1's signal never arrives
        out, _ = embs[0].forward(idx, None, None, 8, 4, CombineMode.kSum)
        torch.cuda.synchronize()
        assert bool(torch.isnan(out).all()), "stale slots were summed"
        assert embs[0].status() == 2  # 1 + the missing rank
        with pytest.raises(CuEmbedError, match="timed out"):
            embs[0].forward(idx, None, None, 8, 4, CombineMode.kSum)
    finally:
        embs[0].set_peer_timeout_ms(60000)
        torch.cuda.synchronize()
        for e in embs:
            e.close()


@pytest.mark.gpu
def test_second_begin_before_finish_is_refused(cuda_lib):
    """The exchange buffers are two epochs deep: a second forward_begin while
    the first is unfinished could overwrite slots a slower peer still reduces
    (ADVICE r1); the object refuses it."""
    import gpu_helpers as gh
    from cuembed_b200 import peer
    from cuembed_b200.api import CuEmbedError
    from cuembed_b200.sharded_p2p import PeerShardedEmbedding
    dev = torch.device(gh.DEV)
    group = peer.LocalPeerGroup(1, dev)
    table = torch.ones(10, 32, dtype=torch.float32, device=dev)
    emb = PeerShardedEmbedding(table, 10, rank=0, world=1,
                               buffers=lambda kind, nbytes: group.alloc(nbytes)[0])
    idx = torch.zeros(8, dtype=torch.int32, device=dev)
    try:
        pend = emb.forward_begin(idx, None, None, 4, 2, CombineMode.kSum)
        with pytest.raises(CuEmbedError, match="in flight"):
            emb.forward_begin(idx, None, None, 4, 2, CombineMode.kSum)
        out, _ = emb.forward_finish(pend)
        torch.cuda.synchronize()
        assert torch.equal(out, torch.full((4, 32), 2.0, device=dev))
    finally:
        emb.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world,dt,it", [(2, F32, np.int32), (4, F16, np.int64)])
def test_virtual_ranks_concat(cuda_lib, oracle, world, dt, it):
    p = Problem(32, 48, 6, "concat", compressed=True, num_categories=301, dt=dt,
                index_dtype=it, seed=67)
    outs, grads, rows, _ = run_virtual(p, world, oracle)[0]
    want = p.cpu_forward(oracle)
    got = np.concatenate([helpers.raw(o) for o in outs])
    assert helpers.bits_equal(got, helpers.raw(want))
    check_backward(p, world, oracle, grads, rows)


# ------------------------------------------------------------- two real GPUs
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _two_gpu_worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from cuembed_b200.sharded_p2p import PeerShardedEmbedding
        p = Problem(256, 128, 24, "mean", csr=True, compressed=True, num_categories=4001,
                    dt=F16, alpha=1.15, seed=71)
        lo, hi = row_range(p.num_categories, world, rank)
        table = torch.from_numpy(p.table[lo:hi].copy()).to(dev)
        # fp32 partials: bit-exact against the rank-ordered restatement below
        emb = PeerShardedEmbedding(table, p.num_categories, partial_dtype=torch.float32)
        idx = torch.from_numpy(p.indices).to(dev)
        off = torch.from_numpy(p.offsets).to(dev)
        per = p.batch // world
        gy = torch.from_numpy(p.grad_y[rank * per:(rank + 1) * per].copy()).to(dev)
        outs = []
        for _ in range(4):  # several epochs: flags and both buffer parities
            out, ctx = emb.forward(idx, off, None, p.batch, 0, CombineMode.kMean)
            grad, rows = emb.backward(gy, ctx, compressed=True)
            torch.cuda.synchronize()
            outs.append((out.cpu().numpy(), grad.cpu().numpy(), rows.cpu().numpy()))
        status = emb.status()
        emb.close()
        results[rank] = (outs, status, (lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpus_ipc_peer_exchange(cuda_lib, oracle):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_two_gpu_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    p = Problem(256, 128, 24, "mean", csr=True, compressed=True, num_categories=4001,
                dt=F16, alpha=1.15, seed=71)
    want = to_f32(helpers.cast_elems(expected_forward(p, world, oracle), F16))
    _, t_idx, t_sid, t_w, remapped = p.cpu_transpose(oracle)
    (g_all, inv_all), _ = p.cpu_backward(oracle, t_idx, t_sid, t_w, remapped, acc_f32=True)
    for step in range(4):
        got = np.concatenate([to_f32(results[r][0][step][0]) for r in range(world)])
        assert np.array_equal(got, want)
        for r in range(world):
            outs, status, (lo, hi) = results[r]
            assert status == 0
            sel = (inv_all >= lo) & (inv_all < hi)
            assert np.array_equal(outs[step][2], inv_all[sel])
            assert helpers.value_equal(outs[step][1], g_all[sel])
