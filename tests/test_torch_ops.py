"""PyTorch custom ops + autograd (cuembed_b200/torch_ops.py).

Follows the reference's own script examples/pytorch/cuembed_test.py:15-182
(forward / backward against nn.EmbeddingBag, inference fast path,
non-contiguous inputs, torch.compile), extended to the dtypes and modes the
new ops add.  CPU part: registration, schemas and fake-tensor shape
propagation (no kernel runs).
"""
import numpy as np
import pytest
import torch
from torch import nn


def _ops():
    from cuembed_b200 import torch_ops
    return torch_ops


def test_ops_registered_with_reference_names_and_schemas():
    _ops()
    ns = torch.ops.cuembed_pyt
    # names and argument order of examples/pytorch/cuembed_embedding.cu:166-182
    s = str(ns.cuembed_embedding_forward.default._schema)
    assert "Tensor params, Tensor indices, Tensor offsets, Tensor? weights" in s and "str mode" in s
    s = str(ns.cuembed_transpose.default._schema)
    assert "Tensor rows, Tensor cols, Tensor? weights" in s and "(Tensor, Tensor, Tensor)" in s
    s = str(ns.cuembed_embedding_backward.default._schema)
    assert ("Tensor y_grad, SymInt num_categories, Tensor transpose_indices, "
            "Tensor transpose_sample_ids, Tensor? transpose_weights") in s
    s = str(ns.cuembed_extract_row_ids_from_csr.default._schema)
    assert "Tensor offsets, SymInt nnz" in s


def test_fake_shape_propagation_without_a_gpu():
    _ops()
    from torch._subclasses.fake_tensor import FakeTensorMode
    ns = torch.ops.cuembed_pyt
    with FakeTensorMode():
        p = torch.empty(100, 16, device="cuda", dtype=torch.float16)
        idx = torch.empty(50, dtype=torch.int32, device="cuda")
        off = torch.empty(11, dtype=torch.int32, device="cuda")
        w = torch.empty(50, dtype=torch.float16, device="cuda")
        out = ns.cuembed_embedding_forward(p, idx, off, w, "sum")
        assert out.shape == (10, 16) and out.dtype == torch.float16
        r = ns.cuembed_extract_row_ids_from_csr(off[:-1], 50)
        assert r.shape == (50,) and r.dtype == torch.int32
        t_r, t_c, t_w = ns.cuembed_transpose(r, idx, w)
        assert t_r.shape == (50,) and t_c.shape == (50,) and t_w.shape == (50,)
        assert ns.cuembed_transpose(r, idx, None)[2].shape == (0,)
        g = ns.cuembed_embedding_backward(out, 100, t_r, t_c, t_w)
        assert g.shape == (100, 16) and g.dtype == torch.float16


def test_cpu_tensors_are_rejected():
    """No CPU implementation is registered: the op must refuse, not fall back."""
    _ops()
    p = torch.zeros(4, 4)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.cuembed_pyt.cuembed_embedding_forward(
            p, torch.zeros(2, dtype=torch.int64), torch.tensor([0, 1, 2]), None, "sum")


# ------------------------------------------------------------------- GPU
def _bag(k, d, mode, dtype):
    return nn.EmbeddingBag(num_embeddings=k, embedding_dim=d, mode=mode,
                           include_last_offset=True, padding_idx=None,
                           dtype=dtype).to(device="cuda")


def _ragged(n_bags, max_len, k, idx_dtype, gen):
    lens = torch.randint(0, max_len + 1, (n_bags,), generator=gen)
    offsets = torch.zeros(n_bags + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(lens, 0)
    nnz = int(offsets[-1])
    idx = torch.randint(0, k, (nnz,), generator=gen)
    return idx.to(idx_dtype).cuda(), offsets.to(idx_dtype).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("weighted", [False, True])
def test_against_embedding_bag_reference_script_shapes(cuda_lib, weighted):
    """examples/pytorch/cuembed_test.py:148-181 (k=2048, d=64, one lookup per bag)."""
    ops = _ops()
    torch.manual_seed(0)
    k, n = 2048, 104217
    eb = _bag(k, 64, "sum", torch.float32)
    indices = (k * torch.rand([n], device="cuda")).to(torch.long)
    offsets = torch.arange(0, n + 1, device="cuda", dtype=torch.long)
    weights = torch.rand([n], device="cuda") if weighted else None
    res = ops.cuemb_embedding(eb.weight, indices, offsets, weights)
    ref = eb(indices, offsets, weights)
    assert torch.equal(res, ref)  # 'fprop test pass'
    eb.weight.grad = None
    torch.mean(res).backward()
    grad_res = eb.weight.grad.clone()
    eb.weight.grad = None
    torch.mean(ref).backward()
    assert torch.allclose(grad_res, eb.weight.grad)  # 'bprop test pass'


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,idt,mode", [
    (torch.float32, torch.int64, "sum"), (torch.float32, torch.int32, "mean"),
    (torch.float16, torch.int64, "sum"), (torch.bfloat16, torch.int32, "sum"),
    (torch.float16, torch.int32, "mean")])
def test_ragged_bags_dtypes_modes(cuda_lib, dtype, idt, mode):
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    k, d = 5000, 128
    eb = _bag(k, d, mode, torch.float32)
    with torch.no_grad():
        eb.weight.copy_(torch.round(eb.weight * 4) / 4)  # exactly representable in bf16
    idx, off = _ragged(700, 40, k, idt, g)
    w16 = eb.weight.detach().to(dtype).requires_grad_(True)
    res = ops.cuemb_embedding(w16, idx, off, None, mode=mode)
    ref = eb(idx.long(), off.long())
    tol = {torch.float32: 1e-6, torch.float16: 2e-3, torch.bfloat16: 2e-2}[dtype]
    assert torch.allclose(res.float(), ref, rtol=tol, atol=tol)
    gy = torch.randint(-3, 4, res.shape, device="cuda").to(dtype)
    res.backward(gy)
    eb.weight.grad = None
    ref.backward(gy.float())
    assert torch.allclose(w16.grad.float(), eb.weight.grad, rtol=tol, atol=tol * 8)


@pytest.mark.gpu
def test_sparse_gradient_matches_dense(cuda_lib):
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    k, d = 30000, 64
    w = torch.randn(k, d, device="cuda", requires_grad=True)
    idx, off = _ragged(512, 30, k, torch.int64, g)
    gy = torch.randn(512, d, device="cuda")
    ops.cuemb_embedding(w, idx, off).backward(gy)
    dense = w.grad.clone()
    w.grad = None
    ops.cuemb_embedding(w, idx, off, sparse_grad=True).backward(gy)
    sp = w.grad
    assert sp.is_sparse and sp._nnz() == int(torch.unique(idx).numel())
    assert torch.equal(sp.to_dense(), dense)


@pytest.mark.gpu
def test_inference_fast_path_and_noncontiguous(cuda_lib):
    """examples/pytorch/cuembed_test.py:36-76."""
    ops = _ops()
    torch.manual_seed(1)
    k, d, n = 958, 32, 20000
    eb = _bag(k, d, "sum", torch.float32)
    indices = torch.randint(0, k, (n,), device="cuda")
    offsets = torch.arange(0, n + 1, device="cuda")
    ref = eb(indices, offsets)
    with torch.no_grad():
        res_nograd = ops.cuemb_embedding(eb.weight, indices, offsets)
    res_frozen = ops.cuemb_embedding(eb.weight.detach(), indices, offsets)
    assert torch.allclose(res_nograd, ref) and torch.allclose(res_frozen, ref)
    assert not res_nograd.requires_grad and not res_frozen.requires_grad
    weight = eb.weight
    w_nc = torch.cat([weight, weight], dim=1).detach()[:, :d]
    idx_nc = torch.stack([indices, indices], dim=1).reshape(-1)[::2]
    assert not w_nc.is_contiguous() and not idx_nc.is_contiguous()
    with torch.no_grad():
        assert torch.allclose(ops.cuemb_embedding(w_nc, idx_nc, offsets), ref)
    grad_mask = torch.ones(ref.shape[0], 2 * d, device="cuda")[:, ::2]
    weight.grad = None
    (ops.cuemb_embedding(weight, idx_nc, offsets) * grad_mask).sum().backward()
    grad_res = weight.grad.clone()
    weight.grad = None
    (eb(indices, offsets) * grad_mask).sum().backward()
    assert torch.allclose(grad_res, weight.grad)


@pytest.mark.gpu
def test_opcheck_and_compile_tracing(cuda_lib):
    """Schema / fake-kernel consistency (torch.library.opcheck) and tracing of
    forward + backward through torch.compile's front end (aot_eager backend:
    the fake registrations are what is under test, not a code generator)."""
    ops = _ops()
    k, d, n = 958, 16, 4096
    w = torch.randn(k, d, device="cuda")
    indices = torch.randint(0, k, (n,), device="cuda")
    offsets = torch.arange(0, n + 1, device="cuda")
    torch.library.opcheck(torch.ops.cuembed_pyt.cuembed_embedding_forward.default,
                          (w, indices, offsets, None, "sum"),
                          test_utils=("test_schema", "test_faketensor"))
    rows = torch.ops.cuembed_pyt.cuembed_extract_row_ids_from_csr(offsets[:-1], n)
    torch.library.opcheck(torch.ops.cuembed_pyt.cuembed_transpose.default,
                          (rows, indices, None),
                          test_utils=("test_schema", "test_faketensor"))

    def run(weight):
        return ops.cuemb_embedding(weight, indices, offsets)

    w_ref = w.clone().requires_grad_(True)
    run(w_ref).sum().backward()
    w_c = w.clone().requires_grad_(True)
    out = torch.compile(run, backend="aot_eager")(w_c)
    out.sum().backward()
    assert torch.allclose(w_ref.grad, w_c.grad, atol=1e-4)
    with torch.no_grad():
        assert torch.allclose(torch.compile(run, backend="aot_eager")(w), run(w))


@pytest.mark.gpu
@pytest.mark.parametrize("optimizer", ["sgd", "adagrad"])
def test_fused_optimizer_step_matches_torch_optim(optimizer):
    """cuemb_embedding_sgd_step (the fused backward + optimizer op, an addition
    over the reference example) against nn.EmbeddingBag + torch.optim on the
    same table: fp32, so one step agrees to rounding (the summation order of the
    gradient differs: <= 1e-5 relative, north_star tolerance)."""
    import torch
    from cuembed_b200.torch_ops import cuemb_embedding, cuemb_embedding_sgd_step
    torch.manual_seed(5)
    dev = "cuda:0"
    rows, width, batch = 2000, 64, 512
    lens = torch.randint(0, 12, (batch,))
    offsets = torch.zeros(batch + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(lens, 0)
    nnz = int(offsets[-1])
    idx = torch.randint(0, rows, (nnz,), dtype=torch.int64)
    table0 = torch.randn(rows, width)
    out_grad = torch.randn(batch, width)
    lr = 0.05

    ref = torch.nn.EmbeddingBag(rows, width, mode="sum", include_last_offset=True).to(dev)
    with torch.no_grad():
        ref.weight.copy_(table0)
    opt = (torch.optim.SGD(ref.parameters(), lr=lr) if optimizer == "sgd" else
           torch.optim.Adagrad(ref.parameters(), lr=lr, eps=1e-10, initial_accumulator_value=0.0))
    ref(idx.to(dev), offsets.to(dev)).backward(out_grad.to(dev))
    opt.step()

    mine = table0.clone().to(dev)
    state = torch.zeros_like(mine) if optimizer == "adagrad" else None
    # the forward of the same table first (the step must not disturb it)
    y = cuemb_embedding(mine, idx.to(dev), offsets.to(dev))
    assert torch.allclose(y, ref_forward(table0.to(dev), idx.to(dev), offsets.to(dev)), rtol=1e-5, atol=1e-5)
    cuemb_embedding_sgd_step(mine, idx.to(dev), offsets.to(dev), out_grad.to(dev), lr,
                             optimizer=optimizer, state=state)
    torch.cuda.synchronize()
    assert torch.allclose(mine, ref.weight.detach(), rtol=1e-5, atol=1e-5)
    untouched = torch.ones(rows, dtype=torch.bool)
    untouched[idx] = False
    assert torch.equal(mine[untouched.to(dev)], table0.to(dev)[untouched.to(dev)])


def ref_forward(table, idx, offsets):
    import torch
    return torch.nn.functional.embedding_bag(idx, table, offsets, mode="sum",
                                             include_last_offset=True)
