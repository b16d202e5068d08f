"""General CNN front-end blocks on the GPU (BatchNorm2d folded, LayerNorm2D incl. the wide-row LayerNorm kernel, residual,
strided convolutions) against the UNMODIFIED reference's outputs (tests/golden/zz_conv_*.npz, gen_golden_conv.py); the same
fixtures are replayed on CPU through the ops' restatements in test_conv_fixtures_cpu below (runs wherever the repository does)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden

CASES = sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLDEN, "zz_conv_*.npz")))


def _run(name, dev, precision):
    from neural_sp_b200.encoders.conv import ConvEncoder
    g = load_golden(name)
    enc = ConvEncoder(**json.loads(str(g["cfg"])))
    enc.load_state_dict({k[3:]: torch.from_numpy(np.asarray(g[k])) for k in g.files if k.startswith("sd.")}, strict=True)
    enc = enc.to(dev).eval()
    enc.set_precision(precision)
    with torch.no_grad():
        ys, ylens = enc(torch.from_numpy(g["xs"]).to(dev), torch.IntTensor(g["xlens"].tolist()))
    assert ylens.tolist() == g["ys_lens"].tolist()
    assert tuple(ys.shape) == g["ys"].shape
    return float(np.abs(ys.float().cpu().numpy() - g["ys"]).max() / np.abs(g["ys"]).max())


@pytest.mark.parametrize("name", CASES)
def test_conv_fixtures_cpu(name, monkeypatch):
    import ops_doubles
    ops_doubles.install(monkeypatch)
    assert _run(name, torch.device("cpu"), "fp32") <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 5e-2)])
@pytest.mark.parametrize("name", CASES)
def test_conv_fixtures_gpu(name, precision, tol):
    assert _run(name, torch.device("cuda:0"), precision) <= tol


@pytest.mark.gpu
@pytest.mark.parametrize("M,D", [(7, 2560), (3, 4100), (5, 2049)])
def test_layernorm_wide_rows(M, D):
    from neural_sp_b200 import ops
    torch.manual_seed(D)
    x = torch.randn(M, D, device="cuda") * 2 + 0.5
    w, b = torch.rand(D, device="cuda") + 0.5, torch.randn(D, device="cuda")
    ref = torch.nn.functional.layer_norm(x, (D,), w, b, 1e-12)
    y, yb = ops.layernorm(x, w, b, 1e-12, out_fp32=True, out_bf16=True)
    assert torch.allclose(y, ref, atol=2e-5, rtol=1e-5)
    assert torch.allclose(yb.float(), ref, atol=3e-2, rtol=2e-2)


@pytest.mark.gpu
def test_mask_rects_kernel_and_staging():
    from neural_sp_b200 import ops
    from neural_sp_b200.frontends.input import pad_and_upload
    from neural_sp_b200.frontends.spec_augment import SpecAugment
    rng = np.random.RandomState(0)
    xs = [rng.randn(n, 80).astype(np.float32) for n in (311, 290, 57)]
    dev, lens = pad_and_upload(xs, "cuda:0")
    assert lens.tolist() == [311, 290, 57] and dev.shape == (3, 311, 80)
    host = dev.cpu()
    for b, x in enumerate(xs):
        assert np.array_equal(host[b, :len(x)].numpy(), x) and float(host[b, len(x):].abs().sum()) == 0
    ref = host.clone()
    fm, tm = [(3, 20), (70, 80)], [(0, 5), (100, 180), (300, 311)]
    for f0, f1 in fm:
        ref[:, :, f0:f1] = 0
    for t0, t1 in tm:
        ref[:, t0:t1] = 0
    out = ops.mask_rects_(dev.clone(), fm, tm)
    assert torch.equal(out.cpu(), ref)
    sa = SpecAugment(27, 100, 2, 2)
    np.random.seed(3)
    o = sa(dev.clone())
    np.random.seed(3)
    f, t = sa.draw(311, 80)
    exp = host.clone()
    for f0, f1 in f:
        exp[:, :, f0:f1] = 0
    for t0, t1 in t:
        exp[:, t0:t1] = 0
    assert torch.equal(o.cpu(), exp)


GRAD_CASES = sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLDEN, "zz_convgrad_*.npz")))


def _run_grads(name, dev, precision):
    """LayerNorm2D front-end in training against the reference's autograd (gen_golden_conv.py main_grads): outputs and every
    parameter gradient.  -> (output error, [(param, error, outliers)])"""
    import sys
    sys.path.insert(0, GOLDEN)
    from gen_golden_conv_weights import loss_weights
    from neural_sp_b200 import autograd as ag
    from neural_sp_b200.encoders.conv import ConvEncoder
    g = load_golden(name)
    enc = ConvEncoder(**json.loads(str(g["cfg"])))
    enc.load_state_dict({k[3:]: torch.from_numpy(np.asarray(g[k])) for k in g.files if k.startswith("sd.")}, strict=True)
    enc = enc.to(dev).train()
    enc.set_precision(precision)
    ys = ag.frontend_forward(enc, torch.from_numpy(g["xs"]).to(dev), 1.0, precision)
    assert tuple(ys.shape) == g["ys"].shape
    e_out = float(np.abs(ys.detach().float().cpu().numpy() - g["ys"]).max() / np.abs(g["ys"]).max())
    w = torch.from_numpy(loss_weights(g["ys"].shape, g["ys_lens"].tolist())).to(dev)
    (ys * w).sum().backward()
    errs = []
    for k, p in enc.named_parameters():
        r = g["g." + k]
        d = np.abs(p.grad.float().cpu().numpy() - r) / max(float(np.abs(r).max()), 1e-12)
        errs.append((k, float(d.max()), int((d > 2e-3).sum())))
    return e_out, errs


@pytest.mark.parametrize("name", GRAD_CASES)
def test_layernorm2d_training_fixtures_cpu(name, monkeypatch):
    import ops_doubles
    from neural_sp_b200 import autograd as ag
    real = ag.frontend_forward
    ops_doubles.install_training(monkeypatch)
    monkeypatch.setattr(ag, "frontend_forward", real)                # the real node over the op restatements
    e_out, errs = _run_grads(name, torch.device("cpu"), "fp32")
    assert e_out <= 1e-5
    # (<= 2 entries per tensor may sit on a ReLU mask bit that rounding flips: see tests/test_frontend_node_cpu.py)
    assert all(n <= 2 and e <= 5e-2 for _, e, n in errs), [x for x in errs if x[2] > 2 or x[1] > 5e-2]


@pytest.mark.gpu
@pytest.mark.parametrize("name", GRAD_CASES)
def test_layernorm2d_training_fixtures_gpu(name):
    e_out, errs = _run_grads(name, torch.device("cuda:0"), "fp32")
    assert e_out <= 1e-4
    assert all(n <= 4 and e <= 5e-2 for _, e, n in errs), [x for x in errs if x[2] > 4 or x[1] > 5e-2]
