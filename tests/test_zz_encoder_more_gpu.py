"""GPU parity of encoder variants added after the round's GPU budget was spent (fixtures: tests/golden/zz_enc_*.npz and
zz_encgrad_*.npz from the UNMODIFIED reference, tests/golden/gen_golden_zz.py): Conformer v2 blocks (bidirectional and
unidirectional) and the plain Transformer with absolute positions (pe_type='add').  Same checks and tolerances as
tests/test_encoder_gpu.py (activations) and tests/test_backward_gpu.py (parameter gradients).
(File name sorts last on purpose; the host wiring of these cases is already pinned on CPU by tests/test_train_wiring_cpu.py.)"""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
from enc_util import build_ours

pytestmark = pytest.mark.gpu
CASES = sorted(os.path.basename(f)[len("zz_enc_"):-4] for f in glob.glob(os.path.join(GOLDEN, "zz_enc_*.npz")))
TOL = {"fp32": 1e-4, "tf32": 5e-3, "bf16": 5e-2}


def _loss_weights(shape, xlens_out, seed=4321):
    w = np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
    for b, n in enumerate(xlens_out):
        w[b, int(n):] = 0.0
    return w


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16"])
@pytest.mark.parametrize("name", CASES)
def test_encoder_matches_reference(name, precision):
    g = load_golden("zz_enc_%s.npz" % name)
    dev = torch.device("cuda:0")
    enc = build_ours(g, dev, precision)
    out = enc(torch.from_numpy(g["xs"]).to(dev), torch.IntTensor(g["xlens"].tolist()), task="all")
    ys, ylens = out["ys"]["xs"], out["ys"]["xlens"]
    assert ylens.tolist() == g["xlens_out"].tolist()
    assert tuple(ys.shape) == g["ys"].shape
    ref, got = g["ys"], ys.float().cpu().numpy()
    if precision == "fp32":
        err = np.abs(got - ref).max() / np.abs(ref).max()
    else:
        err = max(np.abs(got[b, :n] - ref[b, :n]).max() for b, n in enumerate(ylens.tolist())) / np.abs(ref).max()
    assert err <= TOL[precision], (name, precision, err)
    if "ys_sub1" in g.files:
        s1 = out["ys_sub1"]["xs"].float().cpu().numpy()
        assert np.abs(s1 - g["ys_sub1"]).max() / np.abs(g["ys_sub1"]).max() <= TOL[precision] * 2


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-3), ("bf16", 2e-1)])
@pytest.mark.parametrize("name", CASES)
def test_encoder_param_grads_match_reference(name, precision, tol):
    g = load_golden("zz_enc_%s.npz" % name)
    gg = load_golden("zz_encgrad_%s.npz" % name)
    dev = torch.device("cuda:0")
    enc = build_ours(g, dev, precision)
    enc.train()
    out = enc(torch.from_numpy(g["xs"]).to(dev), torch.IntTensor(g["xlens"].tolist()), task="all")
    ys = out["ys"]["xs"]
    assert ys.requires_grad
    w = torch.from_numpy(_loss_weights(tuple(ys.shape), out["ys"]["xlens"].tolist())).to(dev)
    loss = (ys * w).sum()
    if "ys_sub1" in g.files:
        s1 = out["ys_sub1"]["xs"]
        loss = loss + (s1 * torch.from_numpy(_loss_weights(tuple(s1.shape), out["ys_sub1"]["xlens"].tolist(), seed=99)).to(dev)).sum()
    loss.backward()
    gmax = max(float(np.abs(gg[k]).max()) for k in gg.files if k.startswith("g."))
    bad = []
    for k, p in enc.named_parameters():
        ref = torch.from_numpy(gg["g." + k]).double()
        assert p.grad is not None, k
        e = float((p.grad.detach().cpu().double() - ref).abs().max() / max(float(ref.abs().max()), 1e-3 * gmax))
        if not e <= tol:
            bad.append((k, e))
    assert not bad, (name, precision, bad[:10], len(bad))


@pytest.mark.parametrize("mode", ["mean", "drop", "add"])
@pytest.mark.parametrize("B,T,D,f", [(2, 17, 32, 2), (3, 40, 48, 3), (1, 5, 256, 2)])
def test_pool_time_bwd_kernel(mode, B, T, D, f):
    """nsp_pool_time_bwd against torch autograd through the same pooling."""
    from neural_sp_b200 import ops
    torch.manual_seed(T)
    x = torch.randn(B, T, D, device="cuda", requires_grad=True)
    xt = x.transpose(1, 2)
    if mode == "mean":
        ref = torch.nn.functional.avg_pool1d(xt, f, f, ceil_mode=True).transpose(1, 2)
    elif mode == "drop":
        ref = x[:, ::f]
    else:
        Tp = -(-T // f) * f
        ref = torch.nn.functional.pad(xt, (0, Tp - T)).reshape(B, D, Tp // f, f).sum(-1).transpose(1, 2)
    out = ops.pool_time(x.detach(), f, mode)
    assert torch.allclose(out, ref.detach(), atol=1e-6)
    dy = torch.randn_like(out)
    ref.backward(dy)
    dx = ops.pool_time_bwd(dy, T, f, mode)
    assert torch.allclose(dx, x.grad, atol=1e-6)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 3e-2)])
@pytest.mark.parametrize("B,T,d,k,causal", [(3, 45, 64, 7, False), (2, 100, 256, 15, False), (2, 33, 40, 3, True), (1, 5, 512, 31, False)])
def test_conv_module_batchnorm_training_kernels(B, T, d, k, causal, dtype, tol):
    """nsp_dwconv_stats_fwd / nsp_bn_swish_bwd / nsp_dwconv_bwd against torch autograd through
    Swish(BatchNorm_train(depthwise_conv(x))) with statistics over all B*T frames (conformer_convolution.py:113-124)."""
    from neural_sp_b200 import ops
    torch.manual_seed(d + k)
    dev = "cuda"
    x = torch.randn(B, T, d, device=dev).to(dtype)
    w = (torch.randn(d, 1, k, device=dev) * 0.3).requires_grad_(True)
    bias = (torch.randn(d, device=dev) * 0.1).requires_grad_(True)
    g = (torch.rand(d, device=dev) + 0.5).requires_grad_(True)
    bt = (torch.randn(d, device=dev) * 0.1).requires_grad_(True)
    eps = 1e-5
    xr = x.float().clone().requires_grad_(True)
    pad = k - 1 if causal else (k - 1) // 2
    zc = torch.nn.functional.conv1d(torch.nn.functional.pad(xr.transpose(1, 2), (pad, k - 1 - pad)), w, bias, groups=d).transpose(1, 2)
    zf = zc.reshape(-1, d)
    mu, var = zf.mean(0), zf.var(0, unbiased=False)
    u = g * (zc - mu) / torch.sqrt(var + eps) + bt
    y = u * torch.sigmoid(u)
    dy = torch.randn_like(y)
    y.backward(dy)
    taps = w.detach().reshape(d, k).t().contiguous()
    z, stats = ops.dwconv_stats(x, taps, bias.detach(), causal=causal)
    M = B * T
    mean_k = stats[0] / M
    var_k = (stats[1] / M - mean_k * mean_k).clamp_min(0)
    assert float((z.float() - zc.detach()).abs().max()) <= tol * float(zc.abs().max())
    assert float((mean_k - mu.detach()).abs().max()) <= tol and float((var_k - var.detach()).abs().max()) <= tol * max(1.0, float(var.max()))
    yk = ops.conformer_conv(x, taps, bias.detach(), "batch_norm", g.detach(), bt.detach(), eps, mean_k.contiguous(), var_k.contiguous(),
                            causal=causal)
    assert float((yk.float() - y.detach()).abs().max()) <= tol * max(1.0, float(y.abs().max()))
    dz, sums = ops.bn_swish_bwd(z, dy.to(dtype), mean_k.contiguous(), var_k.contiguous(), g.detach(), bt.detach(), eps)
    assert float((sums[0] - bt.grad).abs().max()) <= tol * max(1.0, float(bt.grad.abs().max())) * 4
    assert float((sums[1] - g.grad).abs().max()) <= tol * max(1.0, float(g.grad.abs().max())) * 4
    dtaps, dbias = torch.zeros(k, d, device=dev), torch.zeros(d, device=dev)
    dx = ops.dwconv_bwd(x, taps, dz, dtaps, dbias, causal=causal)
    assert float((dx.float() - xr.grad).abs().max()) <= tol * max(1.0, float(xr.grad.abs().max())) * 4
    assert float((dtaps.t().reshape(d, 1, k) - w.grad).abs().max()) <= tol * max(1.0, float(w.grad.abs().max())) * 8
    assert float((dbias - bias.grad).abs().max()) <= tol * max(1.0, float(bias.grad.abs().max())) * 8


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 5e-2)])
@pytest.mark.parametrize("B,T,d", [(3, 45, 64), (2, 100, 256), (1, 7, 10)])
def test_conv_module_groupnorm_backward_kernel(B, T, d, dtype, tol):
    """nsp_gn2_swish_bwd against torch autograd through Swish(GroupNorm(d/2 groups)) on the per-frame view."""
    from neural_sp_b200 import ops
    torch.manual_seed(d)
    dev = "cuda"
    z = (torch.randn(B, T, d, device=dev) * 1.5).to(dtype)
    zr = z.float().clone().requires_grad_(True)
    g = (torch.rand(d, device=dev) + 0.5).requires_grad_(True)
    bt = (torch.randn(d, device=dev) * 0.1).requires_grad_(True)
    u = torch.nn.functional.group_norm(zr.reshape(B * T, d, 1), d // 2, g, bt, 1e-5).reshape(B, T, d)
    y = u * torch.sigmoid(u)
    dy = torch.randn_like(y)
    y.backward(dy)
    dg, db = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    dz = ops.gn2_swish_bwd(z, dy.to(dtype), g.detach(), bt.detach(), 1e-5, dg, db)
    scale = float(zr.grad.abs().max())
    assert float((dz.float() - zr.grad).abs().max()) <= tol * max(1.0, scale)
    assert float((dg - g.grad).abs().max()) <= tol * max(1.0, float(g.grad.abs().max())) * 4
    assert float((db - bt.grad).abs().max()) <= tol * max(1.0, float(bt.grad.abs().max())) * 4
