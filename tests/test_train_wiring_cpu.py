"""Host logic of the Transformer / Conformer TRAINING path on CPU: the per-block autograd nodes of neural_sp_b200/autograd.py
(which gradient goes where, LayerNorm-fused bias gradients, flat gradient buckets, fused QKV gradient, shared position
projection, LayerDrop / sqrt(d) scaling, hierarchical max-pool, sub-task outputs) run with the library's ops replaced by
torch restatements (tests/ops_doubles.py), and every parameter gradient is compared with the UNMODIFIED reference's
autograd (tests/golden/encgrad_*.npz and zz_encgrad_*.npz).  The CUDA kernels behind the real ops are checked by
tests/test_backward_gpu.py; this test pins what Python does with their results, and runs wherever the repository does."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
from enc_util import build_ours

CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "encgrad_*.npz")) +
               glob.glob(os.path.join(GOLDEN, "zz_encgrad_*.npz")))


def _loss_weights(shape, xlens_out, seed=4321):
    w = np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
    for b, n in enumerate(xlens_out):
        w[b, int(n):] = 0.0
    return w


@pytest.mark.parametrize("case", CASES)
def test_block_training_wiring_matches_reference_gradients(case, monkeypatch):
    import ops_doubles
    ops_doubles.install_training(monkeypatch)
    gg = load_golden(case + ".npz")
    g = load_golden(case.replace("encgrad_", "enc_") + ".npz")
    enc = build_ours(g, torch.device("cpu"), "fp32")
    enc.train()
    out = enc(torch.from_numpy(g["xs"]), torch.IntTensor(g["xlens"].tolist()), task="all")
    ys = out["ys"]["xs"]
    assert ys.requires_grad
    assert out["ys"]["xlens"].tolist() == g["xlens_out"].tolist()
    assert float((ys.detach() - torch.from_numpy(g["ys"])).abs().max()) <= 1e-4 * float(np.abs(g["ys"]).max())
    loss = (ys * torch.from_numpy(_loss_weights(tuple(ys.shape), out["ys"]["xlens"].tolist()))).sum()
    if "ys_sub1" in g.files and case.startswith("zz_"):
        s1 = out["ys_sub1"]["xs"]
        assert float((s1.detach() - torch.from_numpy(g["ys_sub1"])).abs().max()) <= 1e-4 * float(np.abs(g["ys_sub1"]).max())
        loss = loss + (s1 * torch.from_numpy(_loss_weights(tuple(s1.shape), out["ys_sub1"]["xlens"].tolist(), seed=99))).sum()
    assert abs(float(loss.detach()) - float(gg["loss"])) <= 1e-3 * max(1.0, abs(float(gg["loss"])))
    loss.backward()
    gmax = max(float(np.abs(gg[k]).max()) for k in gg.files if k.startswith("g."))
    bad = []
    for k, p in enc.named_parameters():
        if "g." + k not in gg.files:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        ref = torch.from_numpy(gg["g." + k]).double()
        assert p.grad is not None, k
        e = float((p.grad.double() - ref).abs().max() / max(float(ref.abs().max()), 1e-3 * gmax))
        if not e <= 1e-3:
            bad.append((k, e))
    assert not bad, (case, bad[:10], len(bad))
