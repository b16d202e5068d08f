"""Shapes outside the tuned kernels' grid that the reference's test matrix reaches in training (found by the dry run of the
real wrappers, tests/test_dry_wrappers_cpu.py): LayerNorm backward at a width that is not a multiple of 4 (a 10-wide last
projection) or wider than 2048, and the 3x3 weight gradient with three input planes (delta features).  Against torch autograd
of the same op in fp32."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu     # first hardware run: profiles/run_round2_validation.sh, stage 1


@pytest.mark.parametrize("M,D", [(48, 10), (37, 2560), (64, 30)])
def test_layernorm_backward_any_width(M, D):
    from neural_sp_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(M, D, generator=g).to(dev)
    dy = torch.randn(M, D, generator=g).to(dev)
    dres = torch.randn(M, D, generator=g).to(dev)
    gamma = (1 + 0.1 * torch.randn(D, generator=g)).to(dev)
    beta = torch.randn(D, generator=g).to(dev)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-5).backward(dy)
    dgamma, dbeta, dcol = (torch.zeros(D, device=dev) for _ in range(3))
    dx, dxb = ops.layernorm_bwd(dy, x, gamma, 1e-5, dres=dres, dgamma=dgamma, dbeta=dbeta, want_fp32=True, want_bf16=True,
                                dcol=dcol, dcol_alpha=0.5)
    ref = xr.grad + dres
    assert float((dx - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    assert float((dxb.float() - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
    assert float((dgamma - gr.grad).abs().max()) <= 1e-4 * float(gr.grad.abs().max())
    assert float((dbeta - br.grad).abs().max()) <= 1e-4 * float(br.grad.abs().max())
    assert float((dcol - 0.5 * ref.sum(0)).abs().max()) <= 1e-4 * float(ref.sum(0).abs().max())


@pytest.mark.parametrize("CI,CO,chmajor", [(3, 32, True), (3, 32, False), (5, 16, False)])
def test_conv3x3_weight_gradient_odd_input_planes(CI, CO, chmajor):
    from neural_sp_b200 import ops
    dev = torch.device("cuda:0")
    B, T, F = 2, 21, 19
    g = torch.Generator().manual_seed(1)
    a = torch.randn(B, T, F, CI, generator=g)                       # channels-last activations
    dz = torch.randn(B, T, F, CO, generator=g)
    w = torch.zeros(CO, CI, 3, 3, requires_grad=True)
    bias = torch.zeros(CO, requires_grad=True)
    y = torch.nn.functional.conv2d(a.permute(0, 3, 1, 2), w, bias, padding=1)
    y.backward(dz.permute(0, 3, 1, 2))
    a_dev = (a.permute(0, 1, 3, 2).contiguous() if chmajor else a).to(dev)      # [B,T,CI,F] raw-feature layout when chmajor
    dw, db = torch.zeros(CO, CI, 3, 3, device=dev), torch.zeros(CO, device=dev)
    ops.conv3x3_wgrad(a_dev, dz.to(dev), dw, db, B, T, F, in_chmajor=chmajor)
    assert float((dw.cpu() - w.grad).abs().max()) <= 1e-4 * float(w.grad.abs().max())
    assert float((db.cpu() - bias.grad).abs().max()) <= 1e-4 * float(bias.grad.abs().max())


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 3e-2)])
@pytest.mark.parametrize("M,d", [(3 * 37 * 40, 32), (1000, 16), (77, 130)])
def test_batchnorm_backward_without_activation(M, d, dtype, tol):
    """nsp_bn_bwd (BatchNorm2d blocks of the CNN front-end in training) + the k = 1 case of the statistics kernel against torch
    autograd through batch-statistics normalisation."""
    from neural_sp_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + d)
    z = (torch.randn(M, d, generator=g) * 1.5 + 0.3).to(dev).to(dtype)
    du = torch.randn(M, d, generator=g).to(dev).to(dtype)
    gamma = (torch.rand(d, generator=g) + 0.5).to(dev)
    eps = 1e-5
    _, stats = ops.dwconv_stats(z.view(1, M, d), torch.ones(1, d, device=dev), torch.zeros(d, device=dev))
    zf = z.float()
    assert torch.allclose(stats[0], zf.sum(0), rtol=1e-4, atol=1e-2) and torch.allclose(stats[1], (zf * zf).sum(0), rtol=1e-4, atol=1e-2)
    mean, var = zf.mean(0), zf.var(0, unbiased=False)
    zr, gr = zf.clone().requires_grad_(True), gamma.clone().requires_grad_(True)
    b = torch.zeros(d, device=dev, requires_grad=True)
    (gr * (zr - zr.mean(0)) / torch.sqrt(zr.var(0, unbiased=False) + eps) + b).backward(du.float())
    dz, sums = ops.bn_bwd(z, du, mean.contiguous(), var.contiguous(), gamma, eps)
    assert float((dz.float() - zr.grad).abs().max()) <= tol * float(zr.grad.abs().max())
    assert float((sums[0] - b.grad).abs().max()) <= 1e-3 * float(b.grad.abs().max())
    assert float((sums[1] - gr.grad).abs().max()) <= 1e-3 * float(gr.grad.abs().max())
