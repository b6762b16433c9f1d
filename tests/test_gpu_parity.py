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
from helpers import Problem, bits_equal, cast_elems, to_f32, value_equal
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


@pytest.mark.parametrize("case", range(57))
def test_against_cpu_matrix(cuda_lib, oracle, case):
    """Forward, Transpose and Backward on the 57 option sets of
    tests/test_embedding_against_cpu.cu:236-293.  Stricter than the reference's
    own checks (:153-218): everything bit-exact, including weighted forward and
    the order of sample ids / weights inside a row."""
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
