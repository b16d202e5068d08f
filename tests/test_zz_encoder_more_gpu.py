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

# opt-in (NSP_EXPERIMENTAL=1) until the first supervised hardware run at the start of round 2: everything in this file was
# written after the round's GPU budget was spent (host logic pinned on CPU; profiles/run_round2_validation.sh, stage 1)
pytestmark = [pytest.mark.gpu, pytest.mark.experimental]
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
