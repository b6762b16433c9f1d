"""Pins the CPU oracle (oracle/cuembed_oracle.c): against the reference's own
known-answer vectors and against the reference's CPU templates compiled
unchanged into oracle/_ref (when built).  CPU only."""
import numpy as np
import pytest

import helpers
from golden import kat
from helpers import Problem, cast_elems, to_f32, value_equal, bits_equal
from oracle.cpu_lib import BF16, CONCAT, F16, F32, MEAN, SUM, Bf16, CpuLib

DTS = [F32, F16]
ITS = [np.int32, np.int64]


@pytest.fixture(params=["oracle", "ref"])
def lib(request, oracle, reflib):
    return oracle if request.param == "oracle" else reflib


# ------------------------------------------------------------ scalar rounding
def test_half_and_bf16_conversions_match_numpy(oracle):
    rng = np.random.default_rng(0)
    xs = np.concatenate([
        rng.standard_normal(20000).astype(np.float32) * 100,
        rng.standard_normal(20000).astype(np.float32) * 1e-6,
        np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e6, 5.96e-8, 2.98e-8,
                  2.9802322e-08, 6.1e-5, np.inf, -np.inf], np.float32),
    ])
    f = oracle.lib.oracle_f32_to_f16
    f.restype, f.argtypes = __import__("ctypes").c_uint16, [__import__("ctypes").c_float]
    got = np.array([f(float(x)) for x in xs], np.uint16)
    want = xs.astype(np.float16).view(np.uint16)
    assert np.array_equal(got, want)
    g = oracle.lib.oracle_f32_to_bf16
    g.restype, g.argtypes = __import__("ctypes").c_uint16, [__import__("ctypes").c_float]
    gotb = np.array([g(float(x)) for x in xs], np.uint16)
    assert np.array_equal(gotb, Bf16.from_f32(xs).bits)


# ------------------------------------------------------------------ forward
@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("it", ITS)
@pytest.mark.parametrize("csr", [False, True])
def test_forward_kat(lib, dt, it, csr):
    """tests/test_embedding_forward.cu:118-160 (fp16_math = false, :72)."""
    k = kat.FWD
    table = cast_elems(np.array(k["embedding"], np.float32).reshape(5, 4), dt)
    idx = np.array(k["indices"], it)
    off = np.array(k["offsets"], np.int32) if csr else None
    hots = 0 if csr else k["hotness"]
    w = cast_elems(np.array(k["weights"], np.float32), dt)
    cases = [(SUM, None, "sum"), (MEAN, None, "avg"), (SUM, w, "sum_weighted")]
    if not csr:
        cases.append((CONCAT, None, "concat"))
    for mode, ww, key in cases:
        got = lib.forward(table, idx, off, ww, k["batch_size"], hots, mode,
                          embed_width=4)
        assert np.array_equal(to_f32(got).reshape(-1), np.array(k[key], np.float32)), key


def test_forward_argument_checks(oracle):
    k = kat.FWD
    table = np.array(k["embedding"], np.float32).reshape(5, 4)
    idx = np.array(k["indices"], np.int32)
    off = np.array(k["offsets"], np.int32)
    w = np.array(k["weights"], np.float32)
    with pytest.raises(ValueError):  # weights with concat
        oracle.forward(table, idx, None, w, 2, 2, CONCAT, embed_width=4)
    with pytest.raises(ValueError):  # both CSR and fixed
        oracle.forward(table, idx, off, None, 2, 2, SUM, embed_width=4)
    with pytest.raises(ValueError):  # CSR concat
        oracle.forward(table, idx, off, None, 2, 0, CONCAT, embed_width=4)


# ----------------------------------------------------------- index transforms
@pytest.mark.parametrize("it", ITS)
def test_index_transform_kats(lib, it):
    r = kat.README
    assert lib.extract_row_ids_fixed(3, r["fixed_num_hots"], it).tolist() == r["fixed_row_ids"]
    assert lib.extract_row_ids_csr(np.array(r["csr_offsets"], np.int32), 3, it).tolist() == r["csr_row_ids"]
    assert lib.extract_row_ids_concat(4, it).tolist() == r["concat_row_ids"]
    assert lib.compressed_grad_indices(np.array(r["compress_in"], it)).tolist() == r["compress_out"]
    t = kat.TRANSPOSE
    idx = np.array(t["indices"], it)
    for dt in DTS:
        w = cast_elems(np.array(t["weights"], np.float32), dt)
        tr, tc, tw = lib.transpose(np.array(t["sample_ids"], it), idx, w)
        assert tr.tolist() == t["transpose_indices"]
        assert tc.tolist() == t["transpose_sample_ids"]
        assert to_f32(tw).tolist() == t["transpose_weights"]
    tr, tc, _ = lib.transpose(lib.extract_row_ids_concat(4, it), idx, None)
    assert tc.tolist() == t["transpose_sample_ids_concat"]


# ----------------------------------------------------------------- backward
@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("it", ITS)
def test_backward_kat(lib, dt, it):
    """tests/test_embedding_backward.cu:162-202."""
    b = kat.BWD
    ti = np.array(b["transpose_indices"], it)
    tr = np.array(b["transpose_remapped_indices"], it)
    tw = cast_elems(np.array(b["transpose_weights"], np.float32), dt)
    for mode in ("sum", "concat"):
        sid = np.array(b["transpose_sample_ids" + ("_concat" if mode == "concat" else "")], it)
        gy = cast_elems(np.array(b["grad_y_" + mode], np.float32).reshape(-1, 4), dt)
        for weighted in (False, True):
            for compressed in (False, True):
                for skip in (False, True):
                    rows = b["num_unique"] if compressed else b["num_categories"]
                    key = ("cgrad_" if compressed else "grad_") + mode + ("_weighted" if weighted else "")
                    g, inv = lib.backward(gy, 4, rows, ti, sid,
                                          tr if compressed else None,
                                          tw if weighted else None,
                                          skip_grad_init=skip)
                    assert np.array_equal(to_f32(g).reshape(-1), np.array(b[key], np.float32)), key
                    if compressed:
                        assert inv.tolist() == b["inverse_mapping"]


# --------------------------------------- restatement vs reference (oracle/_ref)
def _dtype_cases():
    # (dt, index type, fp16_math) of tests/test_embedding_against_cpu.cu:300-314
    return [(F32, np.int32, False), (F32, np.int64, False), (F16, np.int32, True),
            (F16, np.int64, True), (F16, np.int32, False), (F16, np.int64, False),
            (BF16, np.int32, False), (BF16, np.int32, True)]


@pytest.mark.parametrize("case", range(0, 57, 1))
def test_oracle_equals_reference_on_shape_matrix(oracle, reflib, case):
    """The 57 option sets of tests/test_embedding_against_cpu.cu:236-293; the
    restatement must be BIT-identical to the reference's CPU templates."""
    shape = kat.against_cpu_matrix()[case]
    # Rotate the dtype combos over the matrix so the whole file stays fast.
    combos = _dtype_cases()
    picks = [combos[case % len(combos)], combos[(case + 3) % len(combos)]]
    if shape["batch"] * shape["width"] * shape["hot"] < 1_000_000:
        picks = combos
    for dt, it, fp16_math in picks:
        p = Problem(shape["batch"], shape["width"], shape["hot"], shape["mode"],
                    shape["csr"], shape["weighted"], shape["compressed"], dt=dt,
                    index_dtype=it, seed=100 + case)
        a = p.cpu_forward(oracle, fp16_math=fp16_math)
        b = p.cpu_forward(reflib, fp16_math=fp16_math)
        assert bits_equal(a, b), f"forward {shape} dt={dt}"
        ra = p.cpu_transpose(oracle)
        rb = p.cpu_transpose(reflib)
        for x, y in zip(ra, rb):
            if x is None:
                assert y is None
            else:
                assert bits_equal(x, y), f"transpose {shape} dt={dt}"
        _, t_idx, t_sid, t_w, remapped = ra
        (ga, ia), _ = p.cpu_backward(oracle, t_idx, t_sid, t_w, remapped)
        (gb, ib), _ = p.cpu_backward(reflib, t_idx, t_sid, t_w, remapped)
        assert bits_equal(ga, gb), f"backward {shape} dt={dt}"
        if remapped is not None:
            assert np.array_equal(ia, ib)


def test_oracle_equals_reference_power_law_fp16(oracle, reflib):
    """Hot rows (alpha 1.15): long fp16 accumulation chains with real-valued
    weights, where rounding after every operation matters."""
    p = Problem(256, 64, 32, "sum", weighted=True, compressed=True, dt=F16,
                num_categories=5000, alpha=1.15, seed=5)
    # real-valued weights instead of {0.5, 0.25}
    rng = np.random.default_rng(1)
    p.weights = cast_elems(rng.random(p.nnz).astype(np.float32), F16)
    for fp16_math in (False, True):
        assert bits_equal(p.cpu_forward(oracle, fp16_math=fp16_math),
                          p.cpu_forward(reflib, fp16_math=fp16_math))
    _, t_idx, t_sid, t_w, remapped = p.cpu_transpose(oracle)
    p.grad_y = cast_elems(rng.standard_normal((p.batch, p.width)).astype(np.float32), F16)
    (ga, _), _ = p.cpu_backward(oracle, t_idx, t_sid, t_w, remapped)
    (gb, _), _ = p.cpu_backward(reflib, t_idx, t_sid, t_w, remapped)
    assert bits_equal(ga, gb)


def test_forward_slicing_matches_whole_batch(lib):
    """Batch slicing (used by the multi-threaded CPU baseline) changes nothing."""
    for csr in (False, True):
        p = Problem(101, 36, 7, "sum", csr=csr, weighted=True, dt=F32, seed=3)
        whole = p.cpu_forward(lib)
        parts = np.zeros_like(whole)
        for lo, hi in ((0, 40), (40, 41), (41, 101)):
            lib.forward(p.table, p.indices, p.offsets, p.weights, p.batch,
                        p.num_hots, p.mode, embed_width=p.width, ret=parts,
                        sample_begin=lo, sample_end=hi)
        assert bits_equal(whole, parts)


def test_weighted_mean_extension(oracle):
    """GPU/TF weighted mean (cuembed/include/embedding_lookup_ops.cuh:255-289,
    pinned by tests/test_embedding_ops.cu:281-286): sum * (1 / sum of weights),
    zero vector when the weights sum to zero."""
    table = np.arange(1, 21, dtype=np.float32).reshape(5, 4)
    idx = np.array([1, 3, 0, 4], np.int32)
    w = np.array([1.0, 0.5, 0.0, 0.0], np.float32)
    got = oracle.forward(table, idx, None, w, 2, 2, MEAN, embed_width=4)
    s0 = table[1] * 1.0 + table[3] * 0.5
    want0 = s0 * np.float32(1.0 / 1.5)
    assert np.array_equal(got[0], want0.astype(np.float32))
    assert np.array_equal(got[1], np.zeros(4, np.float32))
