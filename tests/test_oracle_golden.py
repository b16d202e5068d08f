"""Pins the oracle (oracle/*.py) to golden vectors produced by the unmodified reference
(tests/golden/gen_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, split_labels
from oracle import ctc_oracle

CTC_CASES = sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLDEN, "ctc_*.npz")) if "head" not in f)


@pytest.mark.parametrize("name", CTC_CASES)
def test_ctc_oracle_matches_reference(name):
    g = load_golden(name)
    ys = split_labels(g["ys_cat"], g["ylens"])
    loss, grad, nll = ctc_oracle.ctc_forward(g["logits"], ys, g["elens"], float(g["lsm"]))
    assert abs(loss - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    np.testing.assert_allclose(nll, g["nll"], rtol=1e-5, atol=1e-3)
    # the reference's gradient is fp32 (ATen); the oracle is fp64
    np.testing.assert_allclose(grad, g["grad"], rtol=0, atol=2e-4)
    if "trigger_points" in g.files:
        trig = ctc_oracle.forced_align(g["logits"], g["elens"], ys)
        assert np.array_equal(trig, g["trigger_points"])     # bit-exact


def test_ctc_oracle_zero_infinity():
    g = load_golden("ctc_infeasible.npz")
    ys = split_labels(g["ys_cat"], g["ylens"])
    nll, loss, grad = ctc_oracle.ctc_nll_and_grad(g["logits"], ys, g["elens"])
    assert nll[1] == 0.0 and nll[2] == 0.0 and nll[0] > 0       # T < L (+repeats) -> zeroed
    assert np.all(grad[1] == 0) and np.all(grad[2] == 0)


ENC_CASES = sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLDEN, "enc_*.npz")))


@pytest.mark.parametrize("name", ENC_CASES)
def test_encoder_oracle_matches_reference(name):
    import torch
    from enc_util import oracle_cfg
    from oracle import encoder_oracle
    g = load_golden(name)
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd.")}
    out = encoder_oracle.encoder_forward(sd, torch.from_numpy(g["xs"]), g["xlens"].tolist(), oracle_cfg(g), return_layers=True)
    assert out["xlens"] == g["xlens_out"].tolist()
    ref = g["ys"]
    assert np.abs(out["xs"].numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
    for i, a in enumerate(out["layers"]):
        r = g["act.%d" % i]
        assert np.abs(a.numpy() - r).max() <= 1e-5 * np.abs(r).max(), i
    if "ys_sub1" in g.files:
        assert np.abs(out["sub1"].numpy() - g["ys_sub1"]).max() <= 1e-5 * np.abs(g["ys_sub1"]).max()


def test_encoder_state_dict_contract_cpu():
    """Our encoder classes expose exactly the reference's state_dict keys and shapes (no GPU needed)."""
    import torch
    from enc_util import golden_cfg, state_dict_of
    from neural_sp_b200.encoders.conformer import ConformerEncoder
    from neural_sp_b200.encoders.conv import ConvEncoder
    from neural_sp_b200.encoders.transformer import TransformerEncoder
    for name in ENC_CASES:
        g = load_golden(name)
        a, conv, kind = golden_cfg(g)
        a = dict(a)
        a["frontend_conv"] = ConvEncoder(**conv) if conv else None
        if kind == "conformer":
            enc = ConformerEncoder(**a)
        else:
            a.pop("kernel_size"), a.pop("normalization")
            enc = TransformerEncoder(**a)
        ref = state_dict_of(g)
        mine = enc.state_dict()
        assert set(mine) == set(ref), (name, set(mine) ^ set(ref))
        for k in ref:
            assert tuple(mine[k].shape) == tuple(ref[k].shape), (name, k)
        enc.load_state_dict(ref, strict=True)


@pytest.mark.parametrize("name", ["rnnt_small.npz", "rnnt_mid.npz"])
def test_rnnt_oracle_matches_torchaudio_golden(name):
    import torch
    from oracle import rnnt_oracle
    g = load_golden(name)
    lp = torch.from_numpy(g["logits"]).log_softmax(-1).numpy()
    nll, loss, grad = rnnt_oracle.rnnt_nll_and_grad(lp, g["ys"], g["flens"], g["ylens"])
    np.testing.assert_allclose(nll, g["nll"], rtol=1e-5)
    np.testing.assert_allclose(grad, g["grad_log_probs"], atol=1e-4, rtol=0)   # torchaudio is fp32


def test_oracle_gradients_match_reference_autograd():
    """The oracle restatement is differentiable (bench.py's CPU baseline trains through it): its parameter gradients
    equal the UNMODIFIED reference's autograd gradients (tests/golden/encgrad_*.npz) for loss = sum(ys * w)."""
    import glob
    import os

    import torch

    from conftest import GOLDEN, load_golden
    from enc_util import oracle_cfg, state_dict_of
    from oracle import encoder_oracle

    for path in sorted(glob.glob(os.path.join(GOLDEN, "encgrad_*.npz"))):
        name = os.path.basename(path)[len("encgrad_"):-4]
        g, gg = load_golden("enc_%s.npz" % name), load_golden("encgrad_%s.npz" % name)
        sd = {k: (v.clone().float().requires_grad_(True) if v.is_floating_point() and ("g." + k) in gg.files else v)
              for k, v in state_dict_of(g).items()}
        out = encoder_oracle.encoder_forward(sd, torch.from_numpy(g["xs"]), g["xlens"].tolist(), oracle_cfg(g))
        ys = out["xs"]
        w = np.random.default_rng(4321).standard_normal(tuple(ys.shape)).astype(np.float32)
        for b, n in enumerate(out["xlens"]):
            w[b, int(n):] = 0.0
        loss = (ys * torch.from_numpy(w)).sum()
        assert abs(float(loss.detach()) - float(gg["loss"])) <= 1e-3 * max(1.0, abs(float(gg["loss"])))
        loss.backward()
        gmax = max(float(np.abs(gg[k]).max()) for k in gg.files if k.startswith("g."))
        for k, v in sd.items():
            if torch.is_tensor(v) and v.requires_grad:
                ref = gg["g." + k]
                err = float(np.abs(v.grad.numpy() - ref).max()) / max(float(np.abs(ref).max()), 1e-3 * gmax)
                assert err <= 2e-3, (name, k, err)
