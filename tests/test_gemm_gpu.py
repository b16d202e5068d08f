"""tcgen05 GEMM (nsp_linear_fwd) against a plain PyTorch fp32 reference of the same op.

Tolerances (relative to max |ref|): fp32 mode (3xTF32) 2e-5; tf32 single pass 2e-3; bf16 is compared
against the fp32 product of the SAME bf16-rounded operands (only accumulation order differs) 1e-4.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, w, bias, act, glu, residual, alpha):
    y = x.double() @ w.double().t()
    if bias is not None:
        y = y + bias.double()
    if glu:
        n = y.shape[1] // 2
        y = y[:, :n] * torch.sigmoid(y[:, n:])
    elif act == "relu":
        y = torch.relu(y)
    elif act == "swish":
        y = y * torch.sigmoid(y)
    y = alpha * y
    if residual is not None:
        y = y + residual.double()
    return y


def _run(prec, M, N, K, bias=True, act=None, glu=False, residual=False, alpha=1.0, out_dtype=torch.float32):
    from neural_sp_b200 import ops
    torch.manual_seed(M * 7 + N * 3 + K)
    dev = "cuda"
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev) if bias else None
    nout = N // 2 if glu else N
    r = torch.randn(M, nout, device=dev) if residual else None
    if prec == "bf16":
        xr, wr = x.bfloat16().float(), w.bfloat16().float()
    else:
        xr, wr = x, w
    ref = _ref(xr, wr, b, act, glu, r, alpha)
    wp = ops.prepare_weight(w, prec)
    out = ops.linear(x, wp, b, prec=prec, act=act, glu=glu, residual=r, alpha=alpha, out_dtype=out_dtype)
    torch.cuda.synchronize()
    assert out.shape == (M, nout)
    tol = {"fp32": 2e-5, "tf32": 2e-3, "bf16": 1e-4}[prec]
    if out_dtype == torch.bfloat16:
        tol = max(tol, 8e-3)
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= tol, (prec, M, N, K, err)


@pytest.mark.parametrize("prec", ["bf16", "tf32", "fp32"])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 512, 256), (8000, 2048, 512), (1000, 1000, 80),
                                   (77, 40, 1280), (4000, 64, 512), (129, 136, 72)])
def test_linear_shapes(prec, M, N, K):
    _run(prec, M, N, K)


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_linear_epilogues(prec):
    _run(prec, 500, 1024, 256, act="swish")
    _run(prec, 500, 256, 1024, act="relu", bias=False)
    _run(prec, 500, 256, 1024, residual=True, alpha=0.5)
    _run(prec, 333, 1024, 512, glu=True)
    _run(prec, 333, 2 * 72, 64, glu=True)
    _run(prec, 500, 1024, 256, act="swish", out_dtype=torch.bfloat16)


def test_linear_inplace_residual_and_bf16_copy():
    from neural_sp_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(700, 512, device="cuda")
    w = torch.randn(256, 512, device="cuda") / 512 ** 0.5
    res = torch.randn(700, 256, device="cuda")
    ref = res.double() + 0.5 * (x.bfloat16().double() @ w.bfloat16().double().t())
    out, out_b = ops.linear(x, ops.prepare_weight(w, "bf16"), None, prec="bf16", residual=res, alpha=0.5, out=res,
                            out2_bf16=True)
    assert out.data_ptr() == res.data_ptr()
    assert (out.double() - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    assert (out_b.double() - ref).abs().max().item() <= 8e-3 * ref.abs().max().item()
