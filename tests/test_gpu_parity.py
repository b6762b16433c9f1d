"""Parity of the CUDA path (through the C ABI) with the CPU oracle:
known-answer vectors, the reference's randomised shape matrix, edge cases.
All comparisons are bit-exact unless a tolerance is written next to them."""
import numpy as np
import pytest
import torch

import cuembed_b200 as ce
import gpu_helpers as gh
import helpers
from golden import kat
from cuembed_b200 import datagen
from helpers import Problem, bits_equal, cast_elems, raw, to_f32, value_equal
from oracle.cpu_lib import BF16, CONCAT, F16, F32, MEAN, SUM, Bf16

pytestmark = pytest.mark.gpu

DTS = [F32, F16, BF16]
ITS = [np.int32, np.int64]


def _diff(a, b, what):
    fa, fb = to_f32(a).reshape(-1), to_f32(b).reshape(-1)
    bad = np.nonzero(~((fa == fb) | (np.isnan(fa) & np.isnan(fb))))[0]
    return (f"{what}: {bad.size} of {fa.size} elements differ; first at {bad[:5]}: "
            f"gpu={fa[bad[:5]]} cpu={fb[bad[:5]]}")


# ------------------------------------------------------------------ KATs
@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("it", ITS)
@pytest.mark.parametrize("csr", [False, True])
def test_forward_kat(cuda_lib, dt, it, csr):
    """tests/test_embedding_forward.cu:118-160 on the GPU path."""
    k = kat.FWD
    for mode, weighted, key in [("sum", False, "sum"), ("mean", False, "avg"),
                                ("sum", True, "sum_weighted"), ("concat", False, "concat")]:
        if csr and mode == "concat":
            continue
        p = Problem.__new__(Problem)
        p.mode, p.dt, p.width, p.batch = helpers.MODES[mode], dt, 4, 2
        p.table = cast_elems(np.array(k["embedding"], np.float32).reshape(5, 4), dt)
        p.indices = np.array(k["indices"], it)
        p.offsets = np.array(k["offsets"], np.int32) if csr else None
        p.num_hots = 0 if csr else 2
        p.weights = cast_elems(np.array(k["weights"], np.float32), dt) if weighted else None
        p.nnz = 4
        got = gh.gpu_forward(p)
        assert np.array_equal(to_f32(got).reshape(-1), np.array(k[key], np.float32)), key


@pytest.mark.parametrize("it", ITS)
def test_index_transform_kats(cuda_lib, it):
    tt = torch.int32 if it == np.int32 else torch.int64
    r = kat.README
    ids = torch.zeros(9, dtype=tt, device=gh.DEV)
    ce.ExtractRowIdsFromFixed(3, 3, ids)
    assert ids.tolist() == r["fixed_row_ids"]
    ids = torch.zeros(5, dtype=tt, device=gh.DEV)
    ce.ExtractRowIdsFromCSR(torch.tensor(r["csr_offsets"], dtype=torch.int32, device=gh.DEV), 3, ids)
    assert ids.tolist() == r["csr_row_ids"]
    ids = torch.zeros(4, dtype=tt, device=gh.DEV)
    ce.ExtractRowIdsForConcat(4, ids)
    assert ids.tolist() == r["concat_row_ids"]
    src = torch.tensor(r["compress_in"], dtype=tt, device=gh.DEV)
    out = torch.zeros_like(src)
    work = torch.empty(ce.ComputeCompressedGradIndices(src, 7, None, None),
                       dtype=torch.uint8, device=gh.DEV)
    ce.ComputeCompressedGradIndices(src, 7, out, work)
    assert out.tolist() == r["compress_out"]

    t = kat.TRANSPOSE
    idx = torch.tensor(t["indices"], dtype=tt, device=gh.DEV)
    sid = torch.tensor(t["sample_ids"], dtype=tt, device=gh.DEV)
    for wdt in (torch.float32, torch.float16, torch.bfloat16):
        w = torch.tensor(t["weights"], dtype=wdt, device=gh.DEV)
        tr, tc, tw = torch.zeros_like(idx), torch.zeros_like(idx), torch.zeros_like(w)
        # Two-call workspace protocol, tests/test_embedding_transpose.cu:68-89.
        lwork = ce.Transpose(sid, idx, w, 4, None, None, None, None)
        work = torch.empty(lwork, dtype=torch.uint8, device=gh.DEV)
        ce.Transpose(sid, idx, w, 4, tr, tc, tw, work)
        assert tr.tolist() == t["transpose_indices"]
        assert tc.tolist() == t["transpose_sample_ids"]
        assert tw.float().tolist() == t["transpose_weights"]
    sidc = torch.arange(4, dtype=tt, device=gh.DEV)
    tr, tc = torch.zeros_like(idx), torch.zeros_like(idx)
    lwork = ce.Transpose(sidc, idx, None, 4, None, None, None, None)
    work = torch.empty(lwork, dtype=torch.uint8, device=gh.DEV)
    ce.Transpose(sidc, idx, None, 4, tr, tc, None, work)
    assert tc.tolist() == t["transpose_sample_ids_concat"]


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("it", ITS)
def test_backward_kat(cuda_lib, dt, it):
    """tests/test_embedding_backward.cu:162-202, incl. skip_grad_init both ways
    (:250-270) and the explicit-workspace entry point."""
    b = kat.BWD
    tt = torch.int32 if it == np.int32 else torch.int64
    ti = torch.tensor(b["transpose_indices"], dtype=tt, device=gh.DEV)
    tr = torch.tensor(b["transpose_remapped_indices"], dtype=tt, device=gh.DEV)
    tw = gh.to_dev(cast_elems(np.array(b["transpose_weights"], np.float32), dt))
    for mode in ("sum", "concat"):
        sid = torch.tensor(b["transpose_sample_ids" + ("_concat" if mode == "concat" else "")],
                           dtype=tt, device=gh.DEV)
        gy = gh.to_dev(cast_elems(np.array(b["grad_y_" + mode], np.float32).reshape(-1, 4), dt))
        for weighted in (False, True):
            for compressed in (False, True):
                for skip in (False, True):
                    for ws in (False, True):
                        rows = b["num_unique"] if compressed else b["num_categories"]
                        key = ("cgrad_" if compressed else "grad_") + mode + ("_weighted" if weighted else "")
                        grad = torch.full((rows, 4), 0.0 if skip else float("nan"),
                                          dtype=gh.TORCH_DT[dt], device=gh.DEV)
                        inv = torch.full((rows,), -1, dtype=tt, device=gh.DEV) if compressed else None
                        work = None
                        if ws:
                            work = torch.empty(ce.backward_workspace_bytes(gh.TORCH_DT[dt], 4, 4, tt),
                                               dtype=torch.uint8, device=gh.DEV)
                        ce.EmbeddingBackward(gy, 4, rows, 4, ti, sid, tr if compressed else None,
                                             tw if weighted else None, skip, grad, inv, work=work)
                        torch.cuda.synchronize()
                        assert grad.float().cpu().reshape(-1).tolist() == b[key], (key, skip, ws)
                        if compressed:
                            assert inv.tolist() == b["inverse_mapping"]


def test_argument_checks_abort_codes(cuda_lib):
    """Misuse that makes the reference abort (embedding_lookup.cuh:260-267,163)
    raises with the same check text."""
    table = torch.zeros(5, 4, device=gh.DEV)
    idx = torch.zeros(4, dtype=torch.int32, device=gh.DEV)
    off = torch.tensor([0, 2, 4], dtype=torch.int32, device=gh.DEV)
    w = torch.ones(4, device=gh.DEV)
    out = torch.zeros(4, 4, device=gh.DEV)
    with pytest.raises(ce.CuEmbedError, match="kConcat"):
        ce.EmbeddingForward(table, 4, idx, None, w, 2, 2, ce.CombineMode.kConcat, out)
    with pytest.raises(ce.CuEmbedError, match="num_hots"):
        ce.EmbeddingForward(table, 4, idx, off, None, 2, 2, ce.CombineMode.kSum, out)
    with pytest.raises(ce.CuEmbedError, match="kConcat"):
        ce.EmbeddingForward(table, 4, idx, off, None, 2, 0, ce.CombineMode.kConcat, out)
    t16 = torch.zeros(5, 3, dtype=torch.float16, device=gh.DEV)
    with pytest.raises(ce.CuEmbedError, match="bytes_per_row"):
        ce.EmbeddingForward(t16, 3, idx, None, None, 2, 2, ce.CombineMode.kSum,
                            torch.zeros(2, 3, dtype=torch.float16, device=gh.DEV))
    with pytest.raises(ce.CuEmbedError, match="no CPU fallback"):
        ce.EmbeddingForward(table.cpu(), 4, idx, None, None, 2, 2, ce.CombineMode.kSum, out)
    # workspace too small (index_transforms.cuh:126)
    with pytest.raises(ce.CuEmbedError, match="lwork"):
        ce.Transpose(idx, idx, None, 4, idx.clone(), idx.clone(), None,
                     torch.empty(16, dtype=torch.uint8, device=gh.DEV))


# ------------------------------------- randomised matrix vs the CPU oracle
def _combos():
    # tests/test_embedding_against_cpu.cu:300-314 (+ bf16, which the reference lacks)
    return [(F32, np.int32, False), (F32, np.int64, False), (F16, np.int32, True),
            (F16, np.int64, True), (F16, np.int32, False), (F16, np.int64, False),
            (BF16, np.int32, False), (BF16, np.int64, True)]


@pytest.fixture(params=["oracle", "ref"])
def checker(request):
    """The CPU side of the comparison: the C restatement (oracle/) and, one hop
    closer, the reference's own CPU templates compiled from /root/reference
    (oracle/_ref/libcuembed_ref.so, prebuilt; travels with the tree)."""
    return request.getfixturevalue("oracle" if request.param == "oracle" else "reflib")


def test_c1_exact_shape_against_reference_cpu(cuda_lib, reflib):
    """BASELINE.json configs[0] at its own shape: 1 M categories x 32 fp32, batch
    1024, fixed hotness 8, sum (the test_embedding_against_cpu path), checked
    against the reference's CPU templates: forward, transpose, compressed and
    full backward, all bit-exact."""
    for compressed in (True, False):
        p = Problem(1024, 32, 8, "sum", compressed=compressed, num_categories=1_048_576,
                    dt=F32, index_dtype=np.int32, alpha=0.0, seed=31)
        want = p.cpu_forward(reflib)
        got = gh.gpu_forward(p)
        assert bits_equal(got, want), _diff(got, want, "C1 forward")
        rows, t_idx, t_sid, t_w, remapped = gh.gpu_transpose(p)
        c_rows, c_idx, c_sid, c_w, c_remapped = p.cpu_transpose(reflib)
        assert np.array_equal(rows.cpu().numpy(), c_rows)
        assert np.array_equal(t_idx.cpu().numpy(), c_idx)
        assert np.array_equal(t_sid.cpu().numpy(), c_sid)
        if compressed:
            assert np.array_equal(remapped.cpu().numpy(), c_remapped)
        (c_grad, c_inv), _ = p.cpu_backward(reflib, c_idx, c_sid, c_w, c_remapped)
        g_grad, g_inv, _ = gh.gpu_backward(p, t_idx, t_sid, t_w, remapped)
        assert bits_equal(g_grad, c_grad), _diff(g_grad, c_grad, "C1 backward")
        if compressed:
            assert np.array_equal(g_inv, c_inv)


@pytest.mark.parametrize("case", range(57))
def test_against_cpu_matrix(cuda_lib, checker, case):
    """Forward, Transpose and Backward on the 57 option sets of
    tests/test_embedding_against_cpu.cu:236-293, against the C restatement AND
    against the reference's own CPU templates.  Stricter than the reference's
    own checks (:153-218): everything bit-exact, including weighted forward and
    the order of sample ids / weights inside a row."""
    oracle = checker
    shape = kat.against_cpu_matrix()[case]
    combos = _combos()
    big = shape["batch"] * shape["width"] * shape["hot"] >= 1_000_000
    picks = [combos[case % 8], combos[(case + 3) % 8], combos[(case + 6) % 8]] if big else combos
    for dt, it, fp16_math in picks:
        p = Problem(shape["batch"], shape["width"], shape["hot"], shape["mode"],
                    shape["csr"], shape["weighted"], shape["compressed"], dt=dt,
                    index_dtype=it, seed=1000 + case)
        tag = f"{shape} dt={dt} it={it.__name__} fp16_math={fp16_math}"
        want = p.cpu_forward(oracle, fp16_math=fp16_math)
        got = gh.gpu_forward(p, fp16_math=fp16_math)
        assert bits_equal(got, want), _diff(got, want, "forward " + tag)

        rows, t_idx, t_sid, t_w, remapped = gh.gpu_transpose(p)
        c_rows, c_idx, c_sid, c_w, c_remapped = p.cpu_transpose(oracle)
        assert np.array_equal(rows.cpu().numpy(), c_rows), "row ids " + tag
        assert np.array_equal(t_idx.cpu().numpy(), c_idx), "transpose indices " + tag
        assert np.array_equal(t_sid.cpu().numpy(), c_sid), "transpose sample ids " + tag
        if p.weighted:
            assert bits_equal(gh.to_host(t_w), c_w), "transpose weights " + tag
        if p.compressed:
            assert np.array_equal(remapped.cpu().numpy(), c_remapped), "remapped " + tag

        (c_grad, c_inv), _ = p.cpu_backward(oracle, c_idx, c_sid, c_w, c_remapped)
        g_grad, g_inv, _ = gh.gpu_backward(p, t_idx, t_sid, t_w, remapped)
        # integer grad_y x {0.5, 0.25} weights: fp32 accumulation is exact, so
        # the fp16/bf16-accumulating oracle and the GPU agree bit for bit.
        assert value_equal(g_grad, c_grad), _diff(g_grad, c_grad, "backward " + tag)
        if p.compressed:
            assert np.array_equal(g_inv, c_inv), "inverse mapping " + tag


# ------------------------------------------------------------- edge cases
@pytest.mark.parametrize("dt", DTS)
def test_mixed_output_type_and_weighted_mean(cuda_lib, oracle, dt):
    """InputT != OutputT (VecCast, embedding_lookup_ops.cuh:237-240) and the
    GPU/TF weighted mean (:255-289)."""
    p = Problem(257, 40, 9, "mean", csr=True, weighted=True, dt=dt, seed=11)
    rng = np.random.default_rng(3)
    p.weights = cast_elems(rng.random(p.nnz).astype(np.float32), dt)
    for out_dt in (F32, F16, BF16):
        for fp16_math in (False, True):
            want = p.cpu_forward(oracle, fp16_math=fp16_math, out_dt=out_dt)
            got = gh.gpu_forward(p, fp16_math=fp16_math, out_dt=out_dt)
            assert bits_equal(got, want), _diff(got, want, f"dt={dt} out={out_dt} lowp={fp16_math}")


def test_empty_and_ragged_bags(cuda_lib, oracle):
    """Empty bags give zeros for sum and mean (embedding_lookup_ops.cuh:275-277,
    embedding_lookup_cpu.hpp:83-86); batch of all-empty bags; hotness 1."""
    for mode in ("sum", "mean"):
        p = Problem(64, 32, 5, mode, csr=True, dt=F32, seed=2)
        p.offsets = np.zeros(65, np.int32)  # every bag empty
        p.indices = np.zeros(1, np.int32)
        p.nnz = 0
        got = gh.gpu_forward(p)
        assert np.array_equal(got, np.zeros((64, 32), np.float32))
    p = Problem(300, 128, 64, "mean", csr=True, dt=F16, seed=4)
    lens = np.diff(p.offsets)
    assert (lens == 0).any() and lens.max() > 32
    assert bits_equal(gh.gpu_forward(p), p.cpu_forward(oracle))
    p = Problem(1000, 64, 1, "sum", dt=F32, seed=5)
    assert bits_equal(gh.gpu_forward(p), p.cpu_forward(oracle))


@pytest.mark.parametrize("width,dt", [(4096, F32), (8192, F16), (1026, F16), (6, F32), (2, F16)])
def test_wide_and_odd_rows(cuda_lib, oracle, width, dt):
    """Rows up to 16 KB (the reference's limit, embedding_lookup.cuh:213-215)
    and beyond, rows that only allow 4-byte vectors."""
    p = Problem(33, width, 11, "sum", weighted=True, compressed=True, dt=dt,
                num_categories=500, seed=6)
    assert bits_equal(gh.gpu_forward(p), p.cpu_forward(oracle))
    rows, t_idx, t_sid, t_w, remapped = gh.gpu_transpose(p)
    c = p.cpu_transpose(oracle)
    (c_grad, c_inv), _ = p.cpu_backward(oracle, *c[1:])
    g_grad, g_inv, _ = gh.gpu_backward(p, t_idx, t_sid, t_w, remapped)
    assert value_equal(g_grad, c_grad), _diff(g_grad, c_grad, "backward")
    assert np.array_equal(g_inv, c_inv)


@pytest.mark.parametrize("it", ITS)
def test_transpose_general_keys(cuda_lib, oracle, it):
    """Transpose is a general COO transpose: negative keys, keys using all
    bytes, one repeated key, tiny and non-tile-multiple sizes."""
    rng = np.random.default_rng(9)
    tt = torch.int32 if it == np.int32 else torch.int64
    info = np.iinfo(it)
    for nnz, lo, hi in [(1, 0, 10), (2, 0, 2), (4097, 0, 50), (70001, info.min, info.max),
                        (50000, -1000, 1000), (30000, 7, 8), (123457, 0, 10_000_000)]:
        cols = rng.integers(lo, hi, size=nnz, dtype=it)
        rows = np.arange(nnz, dtype=it)
        w = rng.random(nnz).astype(np.float32)
        want = oracle.transpose(rows, cols, w)
        d_rows, d_cols, d_w = gh.to_dev(rows), gh.to_dev(cols), gh.to_dev(w)
        tr, tc, tw = torch.zeros_like(d_cols), torch.zeros_like(d_cols), torch.zeros_like(d_w)
        work = torch.empty(ce.Transpose(d_rows, d_cols, d_w, nnz, None, None, None, None),
                           dtype=torch.uint8, device=gh.DEV)
        ce.Transpose(d_rows, d_cols, d_w, nnz, tr, tc, tw, work)
        torch.cuda.synchronize()
        assert np.array_equal(tr.cpu().numpy(), want[0]), (nnz, lo, hi)
        assert np.array_equal(tc.cpu().numpy(), want[1]), (nnz, lo, hi)
        assert np.array_equal(tw.cpu().numpy(), want[2]), (nnz, lo, hi)


@pytest.mark.parametrize("it", ITS)
@pytest.mark.parametrize("batch,hot,weighted", [(1000, 7, False), (3000, 64, True), (70000, 40, False),
                                                (257, 1, True), (33, 100, False)])
def test_transpose_fixed_equals_row_ids_plus_transpose(cuda_lib, oracle, it, batch, hot, weighted):
    """cuembed_transpose_fixed synthesises the sample ids (position / hotness) in
    the first sort pass: outputs identical to ExtractRowIdsFromFixed + Transpose
    (and to the oracle), for hotness below, at and above the warp width and
    tile sizes that do not divide the hotness."""
    p = Problem(batch, 8, hot, "sum", weighted=weighted, dt=F32, index_dtype=it,
                num_categories=5000, alpha=1.15, seed=87)
    _, c_idx, c_sid, c_w, _ = p.cpu_transpose(oracle)
    idx, w = gh.to_dev(p.indices), gh.to_dev(p.weights)
    t_idx, t_sid = torch.full_like(idx, -1), torch.full_like(idx, -1)
    t_w = torch.zeros_like(w) if w is not None else None
    work = torch.empty(ce.Transpose(idx, idx, w, p.nnz, None, None, None, None),
                       dtype=torch.uint8, device=gh.DEV)
    ce.TransposeFixed(idx, w, batch, hot, t_idx, t_sid, t_w, work)
    torch.cuda.synchronize()
    assert np.array_equal(t_idx.cpu().numpy(), c_idx)
    assert np.array_equal(t_sid.cpu().numpy(), c_sid)
    if weighted:
        assert bits_equal(gh.to_host(t_w), c_w)


@pytest.mark.parametrize("dt", [F32, F16])
def test_backward_long_runs_and_power_law(cuda_lib, oracle, dt):
    """Runs that span many chunks and CTAs (one row hit by every sample), and
    a power-law batch: exercises the head/tail stitching and the fix-up kernel."""
    # every sample hits row 3 plus a few others
    p = Problem(4096, 64, 8, "sum", weighted=True, compressed=True, dt=dt,
                num_categories=300, alpha=1.15, seed=21)
    rows, t_idx, t_sid, t_w, remapped = gh.gpu_transpose(p)
    c = p.cpu_transpose(oracle)
    assert np.array_equal(t_idx.cpu().numpy(), c[1])
    assert np.array_equal(t_sid.cpu().numpy(), c[2])
    # fp32-accumulating checker: exact on integer data, any run length
    (c_grad, c_inv), _ = p.cpu_backward(oracle, *c[1:], acc_f32=True)
    for ws in (False, True):
        g_grad, g_inv, _ = gh.gpu_backward(p, t_idx, t_sid, t_w, remapped, explicit_workspace=ws)
        assert value_equal(g_grad, c_grad), _diff(g_grad, c_grad, f"backward ws={ws}")
        assert np.array_equal(g_inv, c_inv)
    # one single run covering everything
    p2 = Problem(5000, 32, 1, "sum", dt=dt, num_categories=50, seed=22)
    p2.indices[:] = 7
    rows, t_idx, t_sid, t_w, remapped = gh.gpu_transpose(p2)
    c = p2.cpu_transpose(oracle)
    (c_grad, _), _ = p2.cpu_backward(oracle, *c[1:], acc_f32=True)
    g_grad, _, _ = gh.gpu_backward(p2, t_idx, t_sid, t_w, remapped)
    assert value_equal(g_grad, c_grad), _diff(g_grad, c_grad, "single run")


def _backward_explicit_ws(p, t_idx, t_sid, t_w, remapped):
    """EmbeddingBackward through the explicit-workspace entry point, output
    poisoned with NaN first."""
    dt = gh.TORCH_DT[p.dt]
    num_rows = (int(remapped[-1].item()) + 1) if p.compressed else p.num_categories
    grad = torch.full((num_rows, p.width), float("nan"), dtype=dt, device=gh.DEV)
    inv = (torch.full((num_rows,), -1, dtype=t_idx.dtype, device=gh.DEV)
           if p.compressed else None)
    work = torch.empty(ce.backward_workspace_bytes(dt, p.width, p.nnz, t_idx.dtype),
                       dtype=torch.uint8, device=gh.DEV)
    ce.EmbeddingBackward(gh.to_dev(p.grad_y), p.width, num_rows, p.nnz, t_idx, t_sid,
                         remapped, t_w, False, grad, inv, work=work)
    torch.cuda.synchronize()
    return gh.to_host(grad), (inv.cpu().numpy() if inv is not None else None)


HOT_CASES = [
    # width, dtype, weighted, compressed, index type
    (32, F32, False, True, np.int32),     # 128-byte rows: warp walker, 4-byte vectors
    (128, F16, True, True, np.int32),     # 256-byte rows: warp walker, 8-byte vectors
    (256, F16, False, True, np.int32),    # the headline row shape (row copies for single hits)
    (256, BF16, True, False, np.int64),   # full gradient, 64-bit indices
    (128, F32, True, True, np.int64),     # 512-byte fp32 rows
    (512, F16, False, True, np.int32),    # 1 KB rows: two column tiles
    (512, F32, True, True, np.int32),     # 2 KB rows: four column tiles
    (1024, F16, False, False, np.int32),  # 2 KB 16-bit rows
    (48, F16, False, True, np.int32),     # 96-byte rows: generic lane-group walker
    (96, F16, False, True, np.int32),     # 192-byte rows: generic walker, 12 of 16 lanes
]


@pytest.mark.parametrize("case", HOT_CASES, ids=lambda c: f"w{c[0]}-dt{c[1]}-{'w' if c[2] else 'u'}")
def test_backward_hot_rows(cuda_lib, oracle, case):
    """Batches in which a few dozen rows receive thousands of lookups each (runs
    that span dozens of chunks: head / through / tail partials, both fix-up
    levels) next to hundreds of rows hit once, for every row shape of the two
    chunk walkers.  Integer gradients and power-of-two weights: every partial
    sum is exact in fp32, so the result must equal the fp32-accumulating oracle
    whatever the association."""
    width, dt, weighted, compressed, it = case
    p = Problem(12288, width, 12, "sum", weighted=weighted, compressed=compressed, dt=dt,
                num_categories=2500, alpha=1.15, seed=41, index_dtype=it)
    rows, t_idx, t_sid, t_w, remapped = gh.gpu_transpose(p)
    c = p.cpu_transpose(oracle)
    assert np.array_equal(t_sid.cpu().numpy(), c[2])
    (c_grad, c_inv), _ = p.cpu_backward(oracle, *c[1:], acc_f32=True)
    g_grad, g_inv = _backward_explicit_ws(p, t_idx, t_sid, t_w, remapped)
    touched = np.zeros(c_grad.shape[0], bool)
    touched[c[4] if compressed else c[1]] = True
    assert value_equal(raw_rows(g_grad, touched), raw_rows(c_grad, touched)), \
        _diff(raw_rows(g_grad, touched), raw_rows(c_grad, touched), "hot rows")
    if compressed:
        assert np.array_equal(g_inv, c_inv)
    # the drop-in entry point (library-owned scratch) gives the same values
    g2, _, _ = gh.gpu_backward(p, t_idx, t_sid, t_w, remapped)
    assert value_equal(g2, g_grad)


def raw_rows(a, mask):
    if isinstance(a, Bf16):
        return Bf16(a.bits[mask])
    return a[mask]


def test_backward_hot_rows_unsorted_samples_and_concat(cuda_lib, oracle):
    """A caller may pass the sample ids of a run in any order (the reference only
    asks for grouped indices, cuembed/README.md): runs with descending sample
    ids, and concat (sample id = lookup position), give the right sums."""
    p = Problem(12288, 64, 12, "sum", weighted=True, compressed=True, dt=F32,
                num_categories=2500, alpha=1.15, seed=43)
    rows, t_idx, t_sid, t_w, remapped = gh.gpu_transpose(p)
    c = p.cpu_transpose(oracle)
    # reverse the (sample id, weight) order inside every run
    keys = c[1]
    starts = np.flatnonzero(np.r_[True, keys[1:] != keys[:-1]])
    ends = np.r_[starts[1:], keys.size]
    perm = np.concatenate([np.arange(e - 1, s - 1, -1) for s, e in zip(starts, ends)])
    r_sid = np.ascontiguousarray(c[2][perm])
    r_w = np.ascontiguousarray(c[3][perm])
    (c_grad, c_inv), _ = p.cpu_backward(oracle, c[1], r_sid, r_w, c[4], acc_f32=True)
    g_grad, g_inv = _backward_explicit_ws(
        p, t_idx, gh.to_dev(r_sid), gh.to_dev(r_w), remapped)
    assert value_equal(g_grad, c_grad), _diff(g_grad, c_grad, "reversed runs")
    assert np.array_equal(g_inv, c_inv)
    # concat: every lookup is its own "sample"
    pc = Problem(4096, 64, 8, "concat", compressed=True, dt=F16,
                 num_categories=300, alpha=1.15, seed=44)
    rows, t_idx, t_sid, t_w, remapped = gh.gpu_transpose(pc)
    cc = pc.cpu_transpose(oracle)
    (c_grad, c_inv), _ = pc.cpu_backward(oracle, *cc[1:], acc_f32=True)
    g_grad, g_inv = _backward_explicit_ws(pc, t_idx, t_sid, t_w, remapped)
    assert value_equal(g_grad, c_grad), _diff(g_grad, c_grad, "concat")
    assert np.array_equal(g_inv, c_inv)


def test_backward_hot_rows_real_valued_and_deterministic(cuda_lib, oracle):
    """Real-valued gradients with long runs: within 1e-5 of sum|terms| of the
    sequential fp32 oracle (north_star tolerance for a different accumulation
    order across chunks), and bit-identical from run to run."""
    rng = np.random.default_rng(45)
    p = Problem(12288, 256, 12, "sum", weighted=True, compressed=True, dt=F32,
                num_categories=2500, alpha=1.15, seed=46)
    p.grad_y = rng.standard_normal((p.batch, p.width)).astype(np.float32)
    p.weights = rng.random(p.nnz).astype(np.float32)
    rows, t_idx, t_sid, t_w, remapped = gh.gpu_transpose(p)
    c = p.cpu_transpose(oracle)
    (c_grad, _), num_rows = p.cpu_backward(oracle, *c[1:])
    g_grad, _ = _backward_explicit_ws(p, t_idx, t_sid, t_w, remapped)
    l1 = np.zeros((num_rows, p.width))
    np.add.at(l1, c[4], np.abs(p.grad_y.astype(np.float64)[c[2]] *
                               c[3].astype(np.float64)[:, None]))
    rel = np.abs(g_grad.astype(np.float64) - c_grad) / np.maximum(l1, 1e-30)
    assert np.max(rel) <= 1e-5, float(np.max(rel))
    for _ in range(3):
        g2, _ = _backward_explicit_ws(p, t_idx, t_sid, t_w, remapped)
        assert bits_equal(g2, g_grad)


@pytest.mark.parametrize("width,dt,csr,weighted,mode", [
    (256, F16, False, False, "sum"),    # the headline row shape: 512-byte rows
    (64, F32, True, True, "mean"),      # 256-byte fp32 rows, CSR, weighted mean
    (128, BF16, False, True, "sum"),    # 256-byte bf16 rows, weighted
    (64, F16, True, False, "mean"),     # 128-byte rows (4-byte vectors), CSR mean
])
def test_forward_hot_row_cache_is_bit_identical(cuda_lib, oracle, width, dt, csr, weighted, mode):
    """cuembed_forward_hot keeps the listed rows in shared memory; accumulation
    order and arithmetic are those of cuembed_forward, so the output must be
    bit-identical to the oracle for ANY list: the real hot rows of the batch
    (from the transposed indices), an empty list, and a list of unrelated rows
    with duplicates and negative entries."""
    p = Problem(3000, width, 24, mode, csr=csr, weighted=weighted, dt=dt,
                num_categories=6000, alpha=1.15, seed=83)
    want = p.cpu_forward(oracle)
    tdt = gh.TORCH_DT[dt]
    cap = ce.forward_hot_capacity(tdt, width)
    assert cap > 0
    _, t_idx, _, _, _ = gh.gpu_transpose(p)
    hot_rows, hot_count = ce.HotRowsFromSorted(t_idx, p.nnz, 8, cap)
    torch.cuda.synchronize()
    n_hot = int(hot_count.item())
    assert n_hot > 10, "the power-law batch has rows with >= 8 hits"
    # every listed row really has >= 8 hits
    counts = np.bincount(p.indices, minlength=p.num_categories)
    listed = hot_rows[:min(n_hot, cap)].cpu().numpy()
    assert (counts[listed] >= 8).all() and len(set(listed.tolist())) == len(listed)
    rng = np.random.default_rng(5)
    junk = torch.from_numpy(np.r_[rng.integers(0, p.num_categories, 40), [-1, -5],
                                  rng.integers(0, 50, 30)].astype(np.int32)).to(gh.DEV)
    lists = [(hot_rows, hot_count),
             (hot_rows, torch.zeros(1, dtype=torch.int32, device=gh.DEV)),
             (junk, torch.tensor([junk.numel()], dtype=torch.int32, device=gh.DEV))]
    for rows_t, count_t in lists:
        ret = torch.full((p.batch, width), float("nan"), dtype=tdt, device=gh.DEV)
        ce.EmbeddingForwardHot(gh.to_dev(p.table), width, gh.to_dev(p.indices),
                               gh.to_dev(p.offsets), gh.to_dev(p.weights), p.batch,
                               p.num_hots, ce.CombineMode(p.mode), ret, rows_t, count_t)
        torch.cuda.synchronize()
        got = gh.to_host(ret)
        assert bits_equal(got, want), _diff(got, want, f"hot cache, list of {int(count_t.item())}")


@pytest.mark.parametrize("dt", DTS)
def test_forward_multi_table(cuda_lib, oracle, dt):
    """cuembed_forward_multi: 35 tables (two launches of
    <= 32) with different row counts, batch sizes, fixed hotness or CSR bags and
    sum or mean, pooled into one [batch, tables * width] activation matrix.
    Every table's slice must be bit-identical to the oracle's forward of that
    table (the same bar as the single-table call)."""
    n_tables, width, max_batch = 35, 32, 300
    rng = np.random.default_rng(61)
    probs = []
    for t in range(n_tables):
        csr = bool(t % 3 == 1)
        mode = "mean" if t % 4 == 2 else "sum"
        batch = max_batch if t % 5 else int(rng.integers(1, max_batch))
        probs.append(Problem(batch, width, int(rng.integers(1, 20)), mode, csr=csr,
                             weighted=True, dt=dt, index_dtype=np.int64,
                             num_categories=int(rng.integers(50, 5000)), alpha=1.05,
                             seed=100 + t))
    tdt = gh.TORCH_DT[dt]
    out = torch.full((max_batch, n_tables * width), float("nan"), dtype=tdt, device=gh.DEV)
    rets = [out[:p.batch, t * width:(t + 1) * width] for t, p in enumerate(probs)]
    ce.EmbeddingForwardMulti(
        [gh.to_dev(p.table) for p in probs], width,
        [gh.to_dev(p.indices) for p in probs],
        [gh.to_dev(p.offsets.astype(np.int32)) if p.csr else None for p in probs],
        [gh.to_dev(p.weights) for p in probs],
        [p.batch for p in probs], [p.num_hots for p in probs],
        [int(p.mode) for p in probs], rets, out_row_stride=n_tables * width)
    torch.cuda.synchronize()
    for t, p in enumerate(probs):
        if p.mode == MEAN:
            # weighted mean is not defined by the CPU reference
            # (utils/include/embedding_lookup_cpu.hpp:51): compare with the
            # single-table GPU call, which test_mixed_output_type_and_weighted_mean pins
            want = gh.gpu_forward(p)
        else:
            want = p.cpu_forward(oracle)
        got = gh.to_host(rets[t].contiguous())
        assert bits_equal(got, want), _diff(got, want, f"table {t}")
    # rows below a shorter batch stay untouched
    for t, p in enumerate(probs):
        if p.batch < max_batch:
            tail = out[p.batch:, t * width:(t + 1) * width]
            assert bool(torch.isnan(tail.float()).all())


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("opt", ["sgd", "adagrad"])
def test_backward_fused_optimizer_step(cuda_lib, oracle, dt, opt):
    """cuembed_backward_update: the row sums of the backward
    applied to the table in place.  The reference has no such kernel
    (README.md:119 lists it as future work), so the yardstick is the exact row
    sums (integer gradients) pushed through a numpy restatement of the update
    as documented in include/cuembed_b200.h, one rounding per operation, which
    the kernel must match bit for bit; untouched rows must keep their bits."""
    p = Problem(6000, 96, 10, "sum", weighted=True, dt=dt, num_categories=3000,
                alpha=1.15, seed=51)
    rows, t_idx, t_sid, t_w, _ = gh.gpu_transpose(p)
    c = p.cpu_transpose(oracle)
    # exact row sums (integer gradients, power-of-two weights)
    g = np.zeros((p.num_categories, p.width))
    np.add.at(g, c[1], to_f32(p.grad_y).astype(np.float64)[c[2]] *
              to_f32(c[3]).astype(np.float64)[:, None])
    g32 = g.astype(np.float32)
    assert np.array_equal(g32.astype(np.float64), g)
    touched = np.zeros(p.num_categories, bool)
    touched[c[1]] = True
    lr, eps = np.float32(0.01), np.float32(1e-8)
    p32 = to_f32(p.table)
    table = gh.to_dev(p.table)
    state = None
    if opt == "sgd":
        # p (+) round_T(-(lr * g)): one add in the table's type (the L2 atomic
        # unit does it); float64 holds the exact sum of two T values
        upd = cast_elems((-lr) * g32, dt)
        want_t = to_f32(p.table).astype(np.float64) + to_f32(upd).astype(np.float64)
        want_touched = (want_t.astype(np.float16) if dt == F16 else
                        cast_elems(want_t.astype(np.float32), dt))
    else:
        rng = np.random.default_rng(52)
        s0 = rng.random((p.num_categories, p.width), dtype=np.float32)
        state = torch.from_numpy(s0.copy()).to(gh.DEV)
        s1 = s0 + g32 * g32
        want_touched = cast_elems(p32 - (lr * g32) / (np.sqrt(s1) + eps), dt)
    want = want_touched
    ce.EmbeddingBackwardUpdate(gh.to_dev(p.grad_y), p.width, p.nnz, t_idx, t_sid, t_w,
                               ce.OPT_SGD if opt == "sgd" else ce.OPT_ADAGRAD,
                               float(lr), table, state=state, eps=float(eps))
    torch.cuda.synchronize()
    got = gh.to_host(table)
    # untouched rows keep their bits; touched rows match the restatement
    assert bits_equal(raw_rows(got, ~touched), raw_rows(p.table, ~touched))
    assert bits_equal(raw_rows(got, touched), raw_rows(want, touched)), \
        _diff(raw_rows(got, touched), raw_rows(want, touched), f"fused {opt}")
    if state is not None:
        want_s = np.where(touched[:, None], s1, s0)
        assert np.array_equal(state.cpu().numpy(), want_s)


def test_backward_skip_grad_init_leaves_other_rows(cuda_lib, oracle):
    """skip_grad_init: rows without a gradient keep their content; rows with a
    gradient are overwritten (documented in include/cuembed_b200.h)."""
    p = Problem(50, 16, 3, "sum", dt=F32, num_categories=400, seed=23)
    rows, t_idx, t_sid, t_w, remapped = gh.gpu_transpose(p)
    g_grad, _, _ = gh.gpu_backward(p, t_idx, t_sid, t_w, remapped, skip_grad_init=True, prefill=5.0)
    c = p.cpu_transpose(oracle)
    (c_grad, _), _ = p.cpu_backward(oracle, *c[1:])
    touched = np.zeros(p.num_categories, bool)
    touched[p.indices] = True
    assert np.array_equal(g_grad[touched], c_grad[touched])
    assert np.all(g_grad[~touched] == 5.0)


def test_backward_real_valued_tolerance(cuda_lib, oracle):
    """Real-valued gradients.  The GPU sums each run in fp32 in a fixed order
    (sequential inside a chunk, chunk partials in chunk order), which is a
    different association than the oracle's strictly sequential loop, so the
    comparison is against an fp64 sum with an fp32 summation bound of
    (8 sqrt(n) + 2) * 2^-24 * sum|terms| for a run of n terms, plus one output
    rounding (2^-11 relative, 2^-25 absolute in the subnormal range) for fp16.
    The oracle itself (sequential fp32 / fp16 accumulation) is held to the
    same yardstick to show the GPU is not the less accurate of the two."""
    rng = np.random.default_rng(31)
    for dt in (F32, F16):
        p = Problem(2048, 64, 16, "sum", weighted=True, compressed=True, dt=dt,
                    num_categories=2000, alpha=1.15, seed=32)
        p.grad_y = cast_elems(rng.standard_normal((p.batch, p.width)).astype(np.float32), dt)
        p.weights = cast_elems(rng.random(p.nnz).astype(np.float32), dt)
        rows, t_idx, t_sid, t_w, remapped = gh.gpu_transpose(p)
        c = p.cpu_transpose(oracle)
        (c_grad, _), num_rows = p.cpu_backward(oracle, *c[1:])
        g_grad, _, _ = gh.gpu_backward(p, t_idx, t_sid, t_w, remapped)
        # fp64 reference and the L1 norm of the terms, per output element
        terms = (to_f32(p.grad_y).astype(np.float64)[c[2]] *
                 to_f32(c[3]).astype(np.float64)[:, None])
        exact = np.zeros((num_rows, p.width))
        l1 = np.zeros((num_rows, p.width))
        np.add.at(exact, c[4], terms)
        np.add.at(l1, c[4], np.abs(terms))
        out_round = 0.0 if dt == F32 else 2.0 ** -11
        # fp32 summation of n terms: error ~ sqrt(n) * 2^-24 * sum|terms| (random
        # walk), worst case n * 2^-24; allow 8 * sqrt(n) + 2 units.
        n_terms = np.zeros(num_rows)
        np.add.at(n_terms, c[4], 1.0)
        bound = ((8.0 * np.sqrt(n_terms) + 2.0) * 2.0 ** -24)[:, None] * l1 \
            + out_round * np.abs(exact) + (0.0 if dt == F32 else 2.0 ** -25) + 1e-12
        gpu_err = np.abs(to_f32(g_grad).astype(np.float64) - exact)
        assert np.all(gpu_err <= bound), float(np.max(gpu_err / (l1 + 1e-30)))
        if dt == F32:
            cpu_err = np.abs(to_f32(c_grad).astype(np.float64) - exact)
            # north_star: "within 1e-5 relative" where accumulation order differs
            rel = np.abs(to_f32(g_grad).astype(np.float64) - to_f32(c_grad)) / np.maximum(l1, 1e-30)
            assert np.max(rel) <= 1e-5
            assert np.max(gpu_err / (l1 + 1e-30)) <= 4 * max(np.max(cpu_err / (l1 + 1e-30)), 2.0 ** -24)
        # deterministic: a second run is bit-identical
        g2, _, _ = gh.gpu_backward(p, t_idx, t_sid, t_w, remapped)
        assert bits_equal(g_grad, g2)


@pytest.mark.parametrize("it", ITS)
@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("mode,csr,weighted,width,hot,with_cache", [
    ("sum", False, False, 256, 64, True),    # the headline row shape, 2 index rounds per bag
    ("sum", False, True, 64, 100, True),     # 4 rounds per bag: the translation pipeline
    ("mean", True, False, 128, 40, True),    # ragged bags
    ("sum", True, True, 24, 7, False),       # pure remapping inside one table, 4-byte vectors
    ("sum", False, False, 1024, 3, False),   # column tiles
])
def test_forward_mapped_addresser(cuda_lib, oracle, dt, it, mode, csr, weighted, width, hot,
                                  with_cache):
    """cuembed_forward_mapped (the reference's embedding-cache addresser hook,
    cuembed/include/embedding_lookup_kernels.cuh:114-115): about a third of the
    rows are redirected to scattered slots of a second (cache) table -- or to
    other rows of the same table -- and the rest read the backing table.  The
    output must be bit-identical to the oracle's forward on the table that the
    mapping describes."""
    p = Problem(777, width, hot, mode, csr=csr, weighted=weighted, dt=dt, index_dtype=it,
                num_categories=5000, alpha=1.05, seed=55)
    rng = np.random.default_rng(5)
    n = p.num_categories
    cached = rng.random(n) < 0.35
    row_map = np.full(n, -1, dtype=it)
    table_bits = raw(p.table)
    if with_cache:
        n_slots = int(cached.sum()) + 17
        slots = rng.permutation(n_slots)[:int(cached.sum())]
        row_map[cached] = slots.astype(it)
        cache_f32 = datagen.make_table(n_slots, width, seed=99)
        cache = cast_elems(cache_f32, dt)
        effective = table_bits.copy()
        effective[cached] = raw(cache)[slots]
    else:
        targets = rng.integers(0, n, size=int(cached.sum()))
        row_map[cached] = targets.astype(it)
        cache = None
        effective = table_bits.copy()
        effective[cached] = table_bits[targets]
    saved = p.table
    p.table = Bf16(effective) if isinstance(saved, Bf16) else effective
    want = p.cpu_forward(oracle)
    p.table = saved
    ret = torch.full((p.batch, width), float("nan"), dtype=gh.TORCH_DT[dt], device=gh.DEV)
    ce.EmbeddingForwardMapped(gh.to_dev(p.table), width, gh.to_dev(p.indices),
                              gh.to_dev(p.offsets), gh.to_dev(p.weights), p.batch,
                              p.num_hots, ce.CombineMode(p.mode), ret,
                              gh.to_dev(row_map), gh.to_dev(cache))
    torch.cuda.synchronize()
    assert bits_equal(gh.to_host(ret), want)
    # concat and fp16_math have no mapped form: an argument error, not a wrong result
    with pytest.raises(ce.CuEmbedError):
        ce.EmbeddingForwardMapped(gh.to_dev(p.table), width, gh.to_dev(p.indices), None, None,
                                  4, 2, ce.CombineMode.kConcat,
                                  torch.empty(8, width, dtype=gh.TORCH_DT[dt], device=gh.DEV),
                                  gh.to_dev(row_map), None)


@pytest.mark.parametrize("it", ITS)
def test_debug_check_lookup(cuda_lib, it):
    """cuembed_debug_check_lookup: the optional bounds check next to kernels that
    (like the reference's, embedding_lookup_ops.cuh:59) carry none."""
    p = Problem(300, 8, 12, "sum", csr=True, index_dtype=it, num_categories=1000, seed=3)
    idx, off = gh.to_dev(p.indices), gh.to_dev(p.offsets)
    ce.DebugCheckLookup(idx, 1000, off)                      # valid: no error
    ce.DebugCheckLookup(idx, 1000)                           # fixed hotness form
    bad = idx.clone()
    bad[37] = 1000                                           # one past the end
    bad[200] = -5
    with pytest.raises(ce.CuEmbedError, match=r"outside \[0, num_rows\).*lookup: 37"):
        ce.DebugCheckLookup(bad, 1000, off)
    bad_off = off.clone()
    bad_off[11] = bad_off[12] + 1                            # bag 11 ends before it starts
    with pytest.raises(ce.CuEmbedError, match=r"offsets.*bag: 10|offsets.*bag: 11"):
        ce.DebugCheckLookup(idx, 1000, bad_off)
    with pytest.raises(ce.CuEmbedError, match="offsets"):
        ce.DebugCheckLookup(idx[:-3], 1000, off)             # offsets run past nnz
