"""Row-sharded multi-GPU mode.

CPU (gloo, world_size 2): the host logic of cuembed_b200.sharded -- row
partition, reduce-scatter of partial sums, mean by GLOBAL bag length,
all-gather of grad_y, gradients staying on the owner -- with the local stages
played by the CPU oracle (tests/sharded_helpers.py).
GPU: the shard kernels (select / finalize) against numpy, and the whole path at
world_size 1 on NCCL against the single-table oracle.
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
from cuembed_b200.sharded import RowShardedEmbedding, row_range
from helpers import Problem, to_f32
from oracle import cpu_lib
from oracle.cpu_lib import F32


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_row_range_partitions_all_rows():
    for rows, world in ((10, 4), (7, 8), (400_000_000, 8), (1, 1)):
        spans = [row_range(rows, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == rows
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0]


def _gloo_worker(rank, world, port, mode_name, csr, weighted, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sharded_helpers import OracleLocalOps
        p = Problem(64, 16, 6, mode_name, csr=csr, weighted=weighted, compressed=True,
                    num_categories=203, dt=F32, seed=77, integer_table=True)
        lo, hi = row_range(p.num_categories, world, rank)
        table = torch.from_numpy(p.table[lo:hi].copy())
        emb = RowShardedEmbedding(table, p.num_categories, ops=OracleLocalOps())
        idx = torch.from_numpy(p.indices)
        off = torch.from_numpy(p.offsets) if p.offsets is not None else None
        w = torch.from_numpy(p.weights) if p.weights is not None else None
        mode = CombineMode(p.mode)
        out, ctx = emb.forward(idx, off, w, p.batch, p.num_hots, mode)
        per = p.batch // world
        gy = torch.from_numpy(p.grad_y[rank * per:(rank + 1) * per].copy())
        grad, rows = emb.backward(gy, ctx, compressed=True)
        results[rank] = (out.numpy(), grad.numpy(), rows.numpy(), (lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode_name,csr,weighted", [("sum", False, False),
                                                    ("mean", True, False),
                                                    ("sum", True, True)])
def test_sharded_forward_backward_gloo_world2(oracle, mode_name, csr, weighted):
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_gloo_worker, args=(world, _free_port(), mode_name, csr, weighted, results),
             nprocs=world, join=True)
    p = Problem(64, 16, 6, mode_name, csr=csr, weighted=weighted, compressed=True,
                num_categories=203, dt=F32, seed=77, integer_table=True)
    want = p.cpu_forward(oracle)
    got = np.concatenate([results[r][0] for r in range(world)])
    # integer-valued table and power-of-two weights: partial sums are exact, so
    # the reduce-scatter order does not matter; mean multiplies once at the end.
    assert np.array_equal(got, want)
    # backward: each rank's compressed gradient == the global gradient
    # restricted to its rows
    _, t_idx, t_sid, t_w, remapped = p.cpu_transpose(oracle)
    (g_all, inv_all), _ = p.cpu_backward(oracle, t_idx, t_sid, t_w, remapped)
    for r in range(world):
        _, grad, rows, (lo, hi) = results[r]
        sel = (inv_all >= lo) & (inv_all < hi)
        assert np.array_equal(rows, inv_all[sel])
        assert np.array_equal(grad, g_all[sel])


# ------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_shard_select_and_finalize_kernels(cuda_lib):
    import cuembed_b200 as ce
    import gpu_helpers as gh
    rng = np.random.default_rng(5)
    for csr, weighted, it in ((False, False, np.int32), (True, True, np.int64),
                              (True, False, np.int32)):
        p = Problem(257, 8, 9, "mean", csr=csr, weighted=weighted, num_categories=1000,
                    dt=F32, index_dtype=it, seed=41)
        lo, hi = 300, 720
        idx, off, w = gh.to_dev(p.indices), gh.to_dev(p.offsets), gh.to_dev(p.weights)
        from cuembed_b200.sharded import CudaLocalOps
        ops = CudaLocalOps()
        l_off, l_idx, l_w = ops.shard_select(idx, off, w, p.batch, p.num_hots, lo, hi)
        torch.cuda.synchronize()
        from sharded_helpers import OracleLocalOps
        c_off, c_idx, c_w = OracleLocalOps().shard_select(
            torch.from_numpy(p.indices), torch.from_numpy(p.offsets) if csr else None,
            torch.from_numpy(p.weights) if weighted else None, p.batch, p.num_hots, lo, hi)
        n = int(c_off[-1])
        assert np.array_equal(l_off.cpu().numpy(), c_off.numpy())
        assert np.array_equal(l_idx.cpu().numpy()[:n], c_idx.numpy()[:n])
        if weighted:
            assert np.array_equal(l_w.cpu().numpy()[:n], c_w.numpy()[:n])
        # finalize == oracle epilogue
        partial = torch.from_numpy(rng.integers(-50, 50, (64, 8)).astype(np.float32))
        want = OracleLocalOps().finalize(partial, ce.CombineMode.kMean,
                                         torch.from_numpy(p.offsets) if csr else None,
                                         p.num_hots, 128, torch.from_numpy(p.weights) if weighted else None,
                                         torch.float16)
        got = ops.finalize(partial.to(gh.DEV), ce.CombineMode.kMean, off, p.num_hots, 128,
                           w, torch.float16)
        torch.cuda.synchronize()
        assert torch.equal(got.cpu(), want)


@pytest.mark.gpu
def test_two_shards_on_one_gpu_sum_to_the_full_result(cuda_lib, oracle):
    """Partials of the two row ranges add up to the single-table forward, and
    the two local backward passes reproduce the global gradient rows."""
    import cuembed_b200 as ce
    import gpu_helpers as gh
    from cuembed_b200.sharded import CudaLocalOps
    p = Problem(512, 64, 16, "sum", weighted=True, compressed=True, num_categories=5000,
                dt=F32, alpha=1.15, seed=43, integer_table=True)
    ops = CudaLocalOps()
    idx, w = gh.to_dev(p.indices), gh.to_dev(p.weights)
    table = gh.to_dev(p.table)
    gy = gh.to_dev(p.grad_y)
    total = torch.zeros(p.batch, p.width, device=gh.DEV)
    _, t_idx, t_sid, t_w, remapped = p.cpu_transpose(oracle)
    (g_all, inv_all), _ = p.cpu_backward(oracle, t_idx, t_sid, t_w, remapped)
    for rank in range(2):
        lo, hi = row_range(p.num_categories, 2, rank)
        l_off, l_idx, l_w = ops.shard_select(idx, None, w, p.batch, p.num_hots, lo, hi)
        total += ops.pool_partial(table[lo:hi].contiguous(), l_idx, l_off, l_w, p.batch)
        nnz = int(l_off[-1].item())
        grad, inv = ops.local_backward(gy, l_off, l_idx, l_w, p.batch, nnz, hi - lo, True)
        sel = (inv_all >= lo) & (inv_all < hi)
        assert np.array_equal(inv.cpu().numpy() + lo, inv_all[sel])
        assert np.array_equal(grad.cpu().numpy(), g_all[sel])
    assert np.array_equal(total.cpu().numpy(), p.cpu_forward(oracle))


@pytest.mark.gpu
def test_sharded_world1_nccl(cuda_lib, oracle):
    import gpu_helpers as gh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("nccl", rank=0, world_size=1,
                            device_id=torch.device(gh.DEV))
    try:
        p = Problem(256, 128, 12, "mean", csr=True, compressed=True, num_categories=3000,
                    dt=helpers.F16, seed=45)
        emb = RowShardedEmbedding(gh.to_dev(p.table), p.num_categories)
        out, ctx = emb.forward(gh.to_dev(p.indices), gh.to_dev(p.offsets), None, p.batch,
                               p.num_hots, CombineMode.kMean)
        torch.cuda.synchronize()
        assert helpers.bits_equal(gh.to_host(out), p.cpu_forward(oracle))
        grad, rows = emb.backward(gh.to_dev(p.grad_y), ctx, compressed=True)
        _, t_idx, t_sid, t_w, remapped = p.cpu_transpose(oracle)
        (g_all, inv_all), _ = p.cpu_backward(oracle, t_idx, t_sid, t_w, remapped)
        assert np.array_equal(rows.cpu().numpy(), inv_all)
        assert helpers.value_equal(gh.to_host(grad), g_all)
    finally:
        dist.destroy_process_group()
