"""Parity at benchmark scale: the models bench.py times (Conformer-M 12L d256 = BASELINE configs[1], Conformer-L 17L d512 =
configs[2]; CTC head fc512, V = 10000, lsm 0.1) at full width and depth, one training step on a small ragged batch, against
the UNMODIFIED reference's forward, loss and autograd (tests/golden/bench_c{2,3}.npz, gen_golden_benchscale.py).

Bounds (relative to the largest magnitude of the compared tensor):
  fp32 mode (3xTF32 GEMMs, fp32 everything else)  activations / loss / every parameter gradient   1e-3   (north_star)
  bf16 mode (the mode bench.py times)             activations 3e-2, loss 5e-3; per parameter tensor: relative L2 error of
                                                  the sampled gradient entries 1.5e-1, L2 norm 5e-2
The bf16 bounds are what 8-bit mantissas allow through 17 blocks: every GEMM operand carries 2^-9 relative rounding, a
block chains ~10 GEMMs whose errors add in quadrature on a residual stream that is kept in fp32, so ~sqrt(170) * 2^-9 = 2.5e-2
on activations; gradients see the forward error twice (saved activations and the backward GEMMs), and the deepest tensors
(the first CNN layer, 288 weights, each a sum over 32 x 200 x 80 noisy positions) collect it all: measured on B200 the worst
tensor is conv.layers.0.conv1.weight with 9 % / 17 % of max|g| on single entries at 2.3 % / 1.6 % error of its norm
(C2 / C3).  Round 1 accepted 2e-1 element-wise on toy models only."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, GOLDEN

sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu

BOUNDS = {"fp32": dict(act=1e-3, loss=1e-3, gmax=1e-3, gl2=1e-3, gnorm=1e-3),
          "bf16": dict(act=3e-2, loss=5e-3, gmax=2.5e-1, gl2=1.5e-1, gnorm=5e-2)}


def _build(wname, g, prec, dev):
    import bench
    from neural_sp_b200.decoders.ctc import CTC
    from neural_sp_b200.encoders.conformer import ConformerEncoder
    from neural_sp_b200.encoders.conv import ConvEncoder
    w = dict(bench.WORKLOADS[wname])
    a = bench.enc_args(w)
    a["frontend_conv"] = ConvEncoder(**bench.conv_args(w))
    enc = ConformerEncoder(**a)
    sd, head = bench.synth_params(w)
    enc.load_state_dict(sd, strict=True)
    ctc = CTC(eos=2, blank=0, enc_n_units=w["d_model"], vocab=w["vocab"], dropout=0.0, lsm_prob=0.1, fc_list="512")
    ctc.load_state_dict(head, strict=True)
    enc, ctc = enc.to(dev).train(), ctc.to(dev).train()
    enc.set_precision(prec)
    ctc.set_precision(prec)
    return enc, ctc


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("tag", ["bench_c2", "bench_c3"])
def test_training_step_matches_reference_at_benchmark_scale(tag, prec):
    g = np.load(os.path.join(GOLDEN, tag + ".npz"))
    dev = torch.device("cuda")
    enc, ctc = _build(str(g["workload"]), g, prec, dev)
    ys, o = [], 0
    for n in g["ylens"]:
        ys.append(g["ys_cat"][o:o + n].tolist())
        o += int(n)
    out = enc(torch.from_numpy(g["xs"]).to(dev), torch.IntTensor(g["xlens"].tolist()), task='ys')['ys']
    assert out['xlens'].tolist() == g["elens"].tolist()
    loss, _ = ctc(out['xs'], out['xlens'], ys)
    loss.backward()
    torch.cuda.synchronize()
    bd = BOUNDS[prec]
    ref = torch.from_numpy(g["eouts"])
    ours = out['xs'].detach().float().cpu()
    for b, n in enumerate(g["elens"]):                   # valid frames (the reference leaves garbage in padded ones too, but
        e = float((ours[b, :n] - ref[b, :n]).abs().max() / ref[b, :n].abs().max())     # CTC never reads them)
        assert e <= bd["act"], ("encoder output", b, e)
    e = abs(float(loss) - float(g["loss"])) / abs(float(g["loss"]))
    assert e <= bd["loss"], ("loss", float(loss), float(g["loss"]))
    worst = {}
    for pre, mod in (("enc.", enc), ("ctc.", ctc)):
        for k, p in mod.named_parameters():
            assert p.grad is not None, k
            gr = p.grad.detach().float().reshape(-1).cpu().double()
            norm, amax = g["gn." + pre + k]
            if amax == 0:
                assert float(gr.abs().max()) == 0, k
                continue
            n = gr.numel()
            idx = np.unique(np.linspace(0, n - 1, 128).astype(np.int64))
            ref_s = torch.cat([torch.from_numpy(g["gh." + pre + k]).double(), torch.from_numpy(g["gs." + pre + k]).double()])
            our_s = torch.cat([gr[:128], gr[torch.from_numpy(idx)]])
            e_max = float((our_s - ref_s).abs().max()) / amax
            e_l2 = float((our_s - ref_s).norm() / ref_s.norm().clamp_min(1e-30))
            e_norm = abs(float(gr.norm()) - norm) / norm
            worst[pre + k] = (e_max, e_l2, e_norm)
    bad = {k: v for k, v in worst.items() if v[0] > bd["gmax"] or v[1] > bd["gl2"] or v[2] > bd["gnorm"]}
    top = sorted(worst.items(), key=lambda kv: -max(kv[1]))[:5]
    print("%s %s: loss rel err %.2e; worst gradients %s" % (tag, prec, e, top))
    assert not bad, (len(bad), sorted(bad.items(), key=lambda kv: -max(kv[1]))[:8])


def test_smoke_model_gradient_norm_is_the_references():
    """Round 1's smoke() printed a gradient norm of 5e9 for the toy model.  The unmodified reference gives 6.6e9 on the same
    step (bench_smoke.npz): LayerNorm(eps=1e-12) over the all-zero rows the CNN front-end produces for padded frames multiplies
    their gradient by 1/sqrt(eps) = 1e6, and conv.bridge.bias collects it.  Ours must reproduce that number, not hide it."""
    from enc_util import build_ours
    from neural_sp_b200.decoders.ctc import CTC
    g = np.load(os.path.join(GOLDEN, "enc_conformer_small.npz"), allow_pickle=True)
    s = np.load(os.path.join(GOLDEN, "bench_smoke.npz"))
    dev = torch.device("cuda")
    enc = build_ours(g, dev, "fp32").train()
    ctc = CTC(eos=2, blank=0, enc_n_units=int(s["head_w"].shape[1]), vocab=40, lsm_prob=0.1)
    ctc.load_state_dict({"output.weight": torch.from_numpy(s["head_w"]), "output.bias": torch.from_numpy(s["head_b"])})
    ctc = ctc.to(dev).train()
    ctc.set_precision("fp32")
    out = enc(torch.from_numpy(g["xs"]).to(dev), torch.IntTensor(g["xlens"].tolist()), task='ys')['ys']
    loss, _ = ctc(out['xs'], out['xlens'], [[5, 6, 7], [8, 9], [10]])
    loss.backward()
    gn = sum(float(p.grad.double().pow(2).sum()) for p in list(enc.parameters()) + list(ctc.parameters())) ** 0.5
    assert abs(float(loss) - float(s["loss"])) <= 1e-3 * abs(float(s["loss"]))
    assert abs(gn - float(s["grad_norm"])) <= 2e-3 * float(s["grad_norm"]), (gn, float(s["grad_norm"]))
