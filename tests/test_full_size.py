"""The headline workload at FULL size (BASELINE.json configs[1]: 10 M x 256 fp16,
batch 65536, hotness 64, alpha 1.15, compressed gradient) through the C ABI.

The CPU oracle needs minutes at this size, so the checks are properties that
do not depend on size, evaluated with independent torch reductions on the
GPU (test infrastructure, like the oracle):

  * transpose: the output is exactly the (index, sample id) pairs of the input
    in lexicographic order (= a stable sort by index, because sample ids ascend
    in the input) -- compared element by element with a torch.sort of the
    packed pairs; remapped indices = dense rank; inverse_mapping = the distinct
    indices in order;
  * forward on an integer-valued table and backward on integer gradients
    (the reference's own gradient recipe, utils/src/embedding_allocation.cu:234-237):
    every partial sum is an integer below 2^24, so ANY summation order gives
    the same fp32 value and the result must equal torch's gather + sum /
    index_add bit for bit after the single rounding to fp16;
  * linearity of the forward on the real-valued U(-1, 1) table: the fp64 sum of
    all outputs equals the fp64 sum of the looked-up rows to 1e-8 of sum|rows|
    (fp32 output, so only the fp32 accumulation error of 64 terms remains);
  * idempotence: a second backward gives identical bits.
"""
import numpy as np
import pytest
import torch

import cuembed_b200 as ce
from cuembed_b200 import datagen

pytestmark = pytest.mark.gpu

ROWS, WIDTH, BATCH, HOT, ALPHA = 10_000_000, 256, 65536, 64, 1.15
DEV = "cuda:0"


@pytest.fixture(scope="module")
def c2():
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * 2 ** 30:
        pytest.skip("needs ~40 GB of device memory")
    wl = datagen.make_workload(ROWS, WIDTH, BATCH, HOT, alpha=ALPHA, seed=1234)
    g = torch.Generator(device=DEV)
    g.manual_seed(99)
    table = torch.empty(ROWS, WIDTH, dtype=torch.float16, device=DEV)
    for r0 in range(0, ROWS, 1 << 20):
        r1 = min(ROWS, r0 + (1 << 20))
        table[r0:r1] = torch.randint(-8, 9, (r1 - r0, WIDTH), generator=g, device=DEV).half()
    grad_y = torch.randint(-10, 11, (BATCH, WIDTH), generator=g, device=DEV).half()
    indices = torch.from_numpy(wl.indices).to(DEV)
    yield table, indices, grad_y
    del table
    torch.cuda.empty_cache()


def _gathered_sum(table, indices, chunk=4096):
    """sum over each bag of table rows, fp32, in chunks of bags."""
    out = torch.empty(BATCH, WIDTH, dtype=torch.float32, device=DEV)
    idx = indices.view(BATCH, HOT).long()
    for b0 in range(0, BATCH, chunk):
        out[b0:b0 + chunk] = table[idx[b0:b0 + chunk].reshape(-1)].float() \
            .view(-1, HOT, WIDTH).sum(1)
    return out


def test_forward_full_size_exact_on_integer_table(c2):
    table, indices, _ = c2
    out = torch.full((BATCH, WIDTH), float("nan"), dtype=torch.float16, device=DEV)
    ce.EmbeddingForward(table, WIDTH, indices, None, None, BATCH, HOT, ce.CombineMode.kSum, out)
    torch.cuda.synchronize()
    want = _gathered_sum(table, indices)
    assert float(want.abs().max()) <= 2048  # integers exactly representable in fp16
    assert torch.equal(out.float(), want)
    # mean: 1 / 64 is a power of two, still exact
    ce.EmbeddingForward(table, WIDTH, indices, None, None, BATCH, HOT, ce.CombineMode.kMean, out)
    torch.cuda.synchronize()
    assert torch.equal(out.float(), want / HOT)


def test_forward_full_size_linearity_on_real_table(c2):
    _, indices, _ = c2
    g = torch.Generator(device=DEV)
    g.manual_seed(123456)
    rows = 2_000_000  # the checksum needs fp64 row sums: a 2 M-row table, indices folded
    table = (torch.rand(rows, WIDTH, generator=g, device=DEV) * 2 - 1).half()
    idx = (indices.long() % rows).to(indices.dtype)
    out = torch.empty(BATCH, WIDTH, dtype=torch.float32, device=DEV)  # fp32 output: no rounding
    ce.EmbeddingForward(table, WIDTH, idx, None, None, BATCH, HOT, ce.CombineMode.kSum, out)
    torch.cuda.synchronize()
    row_sum = table.double().sum(1)
    row_abs = table.double().abs().sum(1)
    want = row_sum[idx.long()].sum()
    scale = row_abs[idx.long()].sum()
    got = out.double().sum()
    assert abs(float(got - want)) <= 1e-8 * float(scale)


def test_transpose_full_size_is_the_stable_sort(c2):
    _, indices, _ = c2
    nnz = BATCH * HOT
    row_ids = torch.empty(nnz, dtype=torch.int32, device=DEV)
    ce.ExtractRowIdsFromFixed(BATCH, HOT, row_ids)
    assert torch.equal(row_ids, (torch.arange(nnz, device=DEV) // HOT).int())
    t_idx = torch.empty_like(indices)
    t_sid = torch.empty_like(indices)
    remapped = torch.empty_like(indices)
    lwork = max(ce.Transpose(row_ids, indices, None, nnz, None, None, None, None),
                ce.ComputeCompressedGradIndices(indices, nnz, None, None))
    work = torch.empty(lwork, dtype=torch.uint8, device=DEV)
    ce.Transpose(row_ids, indices, None, nnz, t_idx, t_sid, None, work)
    ce.ComputeCompressedGradIndices(t_idx, nnz, remapped, work)
    torch.cuda.synchronize()
    packed_in = indices.long() * BATCH + row_ids.long()
    packed_out = t_idx.long() * BATCH + t_sid.long()
    assert torch.equal(packed_out, torch.sort(packed_in).values)
    uniq, inverse = torch.unique_consecutive(t_idx, return_inverse=True)
    assert torch.equal(remapped.long(), inverse)
    assert int(remapped[-1]) + 1 == uniq.numel()


def test_backward_full_size_exact_and_idempotent(c2):
    _, indices, grad_y = c2
    nnz = BATCH * HOT
    row_ids = torch.empty(nnz, dtype=torch.int32, device=DEV)
    ce.ExtractRowIdsFromFixed(BATCH, HOT, row_ids)
    t_idx = torch.empty_like(indices)
    t_sid = torch.empty_like(indices)
    remapped = torch.empty_like(indices)
    lwork = max(ce.Transpose(row_ids, indices, None, nnz, None, None, None, None),
                ce.ComputeCompressedGradIndices(indices, nnz, None, None))
    work = torch.empty(lwork, dtype=torch.uint8, device=DEV)
    ce.Transpose(row_ids, indices, None, nnz, t_idx, t_sid, None, work)
    ce.ComputeCompressedGradIndices(t_idx, nnz, remapped, work)
    num_unique = int(remapped[-1].item()) + 1
    grad = torch.full((num_unique, WIDTH), float("nan"), dtype=torch.float16, device=DEV)
    inv = torch.full((num_unique,), -1, dtype=torch.int32, device=DEV)
    ce.EmbeddingBackward(grad_y, WIDTH, num_unique, nnz, t_idx, t_sid, remapped, None,
                         True, grad, inv)
    torch.cuda.synchronize()
    # integer sums below 2^24: exact in fp32 in any order, rounded once to fp16
    want = torch.zeros(num_unique, WIDTH, dtype=torch.float32, device=DEV)
    for n0 in range(0, nnz, 1 << 19):
        n1 = min(nnz, n0 + (1 << 19))
        want.index_add_(0, remapped[n0:n1].long(), grad_y[t_sid[n0:n1].long()].float())
    assert float(want.abs().max()) < 2 ** 24
    assert torch.equal(grad, want.half())
    assert torch.equal(inv, torch.unique_consecutive(t_idx))
    grad2 = torch.empty_like(grad)
    ce.EmbeddingBackward(grad_y, WIDTH, num_unique, nnz, t_idx, t_sid, remapped, None,
                         True, grad2, inv)
    torch.cuda.synchronize()
    assert torch.equal(grad2.view(torch.int16), grad.view(torch.int16))
    # the fused SGD step on the same sums: p - lr * g with lr = 2^-6 (exact products)
    table = torch.zeros(ROWS // 10, WIDTH, dtype=torch.float32, device=DEV)
    t_idx_small = (t_idx.long() % (ROWS // 10)).int()
    # folding the indices breaks the grouping: sort again for this part
    order = torch.sort(t_idx_small.long() * BATCH + t_sid.long()).indices
    ti, ts = t_idx_small[order].contiguous(), t_sid[order].contiguous()
    ce.EmbeddingBackwardUpdate(grad_y.float(), WIDTH, nnz, ti, ts, None, ce.OPT_SGD,
                               2.0 ** -6, table)
    torch.cuda.synchronize()
    want_t = torch.zeros_like(table)
    for n0 in range(0, nnz, 1 << 19):
        n1 = min(nnz, n0 + (1 << 19))
        want_t.index_add_(0, ti[n0:n1].long(), grad_y[ts[n0:n1].long()].float())
    assert torch.equal(table, -(2.0 ** -6) * want_t)


def test_backward_full_gradient_beyond_2_31_elements(c2):
    """Full (uncompressed) gradient at the headline shape: 10 M x 256 fp16 =
    2.56 G elements (5.12 GB), i.e. element offsets beyond 2^31.  The reference
    addresses gradient rows with `int` arithmetic
    (cuembed/include/embedding_lookup_ops.cuh:610-618) and overflows here; this
    library uses 64-bit byte offsets.  Integer gradients: every sum is exact in
    fp32, so the result must equal torch's index_add bit for bit, rows that no
    lookup touches must be zero (skip_grad_init = false) and a poisoned buffer
    must keep its poison there (skip_grad_init = true)."""
    _, indices, grad_y = c2
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * 2 ** 30:
        pytest.skip("needs ~40 GB of device memory")
    nnz = BATCH * HOT
    assert ROWS * WIDTH >= 2 ** 31
    row_ids = torch.empty(nnz, dtype=torch.int32, device=DEV)
    ce.ExtractRowIdsFromFixed(BATCH, HOT, row_ids)
    t_idx = torch.empty_like(indices)
    t_sid = torch.empty_like(indices)
    work = torch.empty(ce.Transpose(row_ids, indices, None, nnz, None, None, None, None),
                       dtype=torch.uint8, device=DEV)
    ce.Transpose(row_ids, indices, None, nnz, t_idx, t_sid, None, work)
    grad = torch.full((ROWS, WIDTH), float("nan"), dtype=torch.float16, device=DEV)
    ce.EmbeddingBackward(grad_y, WIDTH, ROWS, nnz, t_idx, t_sid, None, None, False, grad, None)
    torch.cuda.synchronize()
    want = torch.zeros(ROWS, WIDTH, dtype=torch.float32, device=DEV)
    for n0 in range(0, nnz, 1 << 19):
        n1 = min(nnz, n0 + (1 << 19))
        want.index_add_(0, t_idx[n0:n1].long(), grad_y[t_sid[n0:n1].long()].float())
    assert float(want.abs().max()) < 2 ** 24
    want16 = want.half()
    del want
    # rows in the upper half of the table lie beyond element offset 2^31
    touched_high = int((t_idx.long() * WIDTH >= 2 ** 31).sum())
    assert touched_high > 100_000  # (hot rows sit at random places of the table)
    for r0 in range(0, ROWS, 1 << 21):  # chunked compare keeps the peak memory low
        r1 = min(ROWS, r0 + (1 << 21))
        assert torch.equal(grad[r0:r1], want16[r0:r1]), f"rows {r0}..{r1}"
    # skip_grad_init: untouched rows keep their previous contents
    grad.fill_(7.0)
    ce.EmbeddingBackward(grad_y, WIDTH, ROWS, nnz, t_idx, t_sid, None, None, True, grad, None)
    torch.cuda.synchronize()
    touched = torch.zeros(ROWS, dtype=torch.bool, device=DEV)
    touched[t_idx.long()] = True
    for r0 in range(0, ROWS, 1 << 21):
        r1 = min(ROWS, r0 + (1 << 21))
        t = touched[r0:r1, None]
        assert torch.equal(grad[r0:r1], torch.where(t, want16[r0:r1], torch.full_like(want16[r0:r1], 7.0)))
    del grad, want16
    torch.cuda.empty_cache()


# ---------------------------------------------------------------- C3 shapes
def test_c3_csr_weighted_int64_full_size_exact():
    """BASELINE.json configs[2]: CSR bags of U{0..64} lookups (mean 32), weighted
    sum and mean, width 128, fp32 and bf16, int64 indices, batch 131072 -- on an
    integer-valued table with the reference's {0.5, 0.25} weights every product
    and partial sum is exact, so forward, transpose (weights travel with their
    pairs) and backward must equal torch's segment sums bit for bit."""
    rows, width, batch, hot = 10_000_000, 128, 131072, 64
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * 2 ** 30:
        pytest.skip("needs ~40 GB of device memory")
    wl = datagen.make_workload(rows, width, batch, hot, alpha=ALPHA, csr=True, weighted=True,
                               index_dtype=np.int64, offset_dtype=np.int32, seed=77)
    nnz = wl.nnz
    indices = torch.from_numpy(wl.indices).to(DEV)
    offsets = torch.from_numpy(wl.offsets).to(DEV)
    weights32 = torch.from_numpy(wl.weights).to(DEV)
    g = torch.Generator(device=DEV)
    g.manual_seed(7)
    lens = (offsets[1:] - offsets[:-1]).long()
    bag_of = torch.repeat_interleave(torch.arange(batch, device=DEV), lens)
    for dt in (torch.float32, torch.bfloat16):
        table = torch.empty(rows, width, dtype=dt, device=DEV)
        for r0 in range(0, rows, 1 << 21):
            r1 = min(rows, r0 + (1 << 21))
            table[r0:r1] = torch.randint(-4, 5, (r1 - r0, width), generator=g, device=DEV).to(dt)
        w = weights32.to(dt)
        # forward, weighted sum
        out = torch.full((batch, width), float("nan"), dtype=dt, device=DEV)
        ce.EmbeddingForward(table, width, indices, offsets, w, batch, 0, ce.CombineMode.kSum, out)
        torch.cuda.synchronize()
        want = torch.zeros(batch, width, dtype=torch.float32, device=DEV)
        for n0 in range(0, nnz, 1 << 19):
            n1 = min(nnz, n0 + (1 << 19))
            want.index_add_(0, bag_of[n0:n1],
                            table[indices[n0:n1]].float() * weights32[n0:n1, None])
        assert float(want.abs().max()) <= 128  # exact in bf16 too (multiples of 1/4)
        assert torch.equal(out.float(), want)
        # unweighted mean: sum * (1 / len), zero vector for empty bags
        ce.EmbeddingForward(table, width, indices, offsets, None, batch, 0,
                            ce.CombineMode.kMean, out)
        torch.cuda.synchronize()
        plain = torch.zeros(batch, width, dtype=torch.float32, device=DEV)
        for n0 in range(0, nnz, 1 << 19):
            n1 = min(nnz, n0 + (1 << 19))
            plain.index_add_(0, bag_of[n0:n1], table[indices[n0:n1]].float())
        inv_len = torch.where(lens > 0, 1.0 / lens.float(), torch.zeros_like(lens, dtype=torch.float32))
        assert torch.equal(out.float(), (plain * inv_len[:, None]).to(dt).float())
        # transpose with weights + backward (weighted, compressed)
        row_ids = torch.empty(nnz, dtype=torch.int64, device=DEV)
        ce.ExtractRowIdsFromCSR(offsets, batch, row_ids)
        assert torch.equal(row_ids, bag_of)
        t_idx, t_sid, t_w = (torch.empty_like(indices), torch.empty_like(indices),
                             torch.empty_like(w))
        remapped = torch.empty_like(indices)
        lwork = max(ce.Transpose(row_ids, indices, w, nnz, None, None, None, None),
                    ce.ComputeCompressedGradIndices(indices, nnz, None, None))
        work = torch.empty(lwork, dtype=torch.uint8, device=DEV)
        ce.Transpose(row_ids, indices, w, nnz, t_idx, t_sid, t_w, work)
        ce.ComputeCompressedGradIndices(t_idx, nnz, remapped, work)
        torch.cuda.synchronize()
        order = torch.sort(indices * batch + row_ids).indices
        assert torch.equal(t_idx, indices[order]) and torch.equal(t_sid, row_ids[order])
        assert torch.equal(t_w.float(), weights32[order])
        num_unique = int(remapped[-1].item()) + 1
        grad_y = torch.randint(-10, 11, (batch, width), generator=g, device=DEV).to(dt)
        grad = torch.full((num_unique, width), float("nan"), dtype=dt, device=DEV)
        inv = torch.empty(num_unique, dtype=torch.int64, device=DEV)
        ce.EmbeddingBackward(grad_y, width, num_unique, nnz, t_idx, t_sid, remapped, t_w,
                             True, grad, inv)
        torch.cuda.synchronize()
        want_g = torch.zeros(num_unique, width, dtype=torch.float32, device=DEV)
        for n0 in range(0, nnz, 1 << 19):
            n1 = min(nnz, n0 + (1 << 19))
            want_g.index_add_(0, remapped[n0:n1],
                              grad_y[t_sid[n0:n1]].float() * t_w[n0:n1, None].float())
        assert float(want_g.abs().max()) < 2 ** 22  # multiples of 1/4 below 2^24 / 4
        assert torch.equal(grad.float(), want_g.to(dt).float())
        assert torch.equal(inv, torch.unique_consecutive(t_idx))
        del table, out, want, plain, grad, want_g
        torch.cuda.empty_cache()
