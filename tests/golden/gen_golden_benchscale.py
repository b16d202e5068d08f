#!/usr/bin/env python3
"""Benchmark-scale parity fixtures from the UNMODIFIED reference (run in the build container; see gen_golden.py).

    python tests/golden/gen_golden_benchscale.py

bench_c2.npz / bench_c3.npz: the models bench.py times (BASELINE configs[1] Conformer-M 12L d256 and configs[2] Conformer-L
17L d512, CTC head fc512, V = 10000, lsm 0.1) at FULL width and depth on a small batch (B = 2, T = 200, ragged), weights =
bench.synth_params (the tensors every bench arm loads).  One training step of the reference: ConformerEncoder.forward ->
CTC.forward -> backward.  Stored: the inputs, the encoder output, the loss and, for every parameter of encoder + head, a
summary of d loss / d param that fits a fixture: its L2 norm, max |g|, the first 128 entries and 128 evenly strided entries
(Conformer-L has 109 M parameters; the full gradient would be 436 MB).

bench_smoke.npz: the same step for __graft_entry__.smoke()'s toy model (enc_conformer_small weights + a seeded CTC head,
V = 40): loss and the global gradient norm -- round 1's smoke printed 5e9 and nobody knew whether the reference agreed.
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402

import_reference()
import bench  # noqa: E402

NSAMP = 128


def grad_summary(g):
    g = g.detach().reshape(-1).double()
    n = g.numel()
    idx = np.unique(np.linspace(0, n - 1, NSAMP).astype(np.int64))
    return dict(norm=float(g.norm()), amax=float(g.abs().max()), head=g[:NSAMP].float().numpy(),
                strided=g[torch.from_numpy(idx)].float().numpy())


def reference_step(enc, ctc, xs, xlens, ys):
    for p in list(enc.parameters()) + list(ctc.parameters()):
        p.grad = None
    out = enc(torch.from_numpy(xs), torch.IntTensor(xlens), task='ys')['ys']
    loss, _ = ctc(out['xs'], out['xlens'], ys)
    loss.backward()
    return out, loss


def gen_bench(tag, wname, B=2, T=200, xlens=(200, 170)):
    w = dict(bench.WORKLOADS[wname], B=B, T=T)
    conv_mod = importlib.import_module('neural_sp.models.seq2seq.encoders.conv')
    conf_mod = importlib.import_module('neural_sp.models.seq2seq.encoders.conformer')
    ctc_mod = importlib.import_module('neural_sp.models.seq2seq.decoders.ctc')
    a = bench.enc_args(w)
    a["frontend_conv"] = conv_mod.ConvEncoder(**bench.conv_args(w))
    enc = conf_mod.ConformerEncoder(**a)
    sd, head = bench.synth_params(w)
    enc.load_state_dict(sd, strict=True)
    ctc = ctc_mod.CTC(eos=2, blank=0, enc_n_units=w["d_model"], vocab=w["vocab"], dropout=0.0, lsm_prob=0.1, fc_list="512")
    ctc.load_state_dict(head, strict=True)
    enc.train(), ctc.train()                         # dropouts are 0
    rng = np.random.default_rng(4321)
    xs = np.zeros((B, T, 80), np.float32)
    for b, n in enumerate(xlens):
        xs[b, :n] = rng.standard_normal((n, 80)).astype(np.float32)
    ys = [rng.integers(4, w["vocab"], size=max(1, int(0.45 * n / 8))).tolist() for n in xlens]
    out, loss = reference_step(enc, ctc, xs, list(xlens), ys)
    save = dict(xs=xs, xlens=np.array(xlens, np.int32), ys_cat=np.array([v for y in ys for v in y], np.int32),
                ylens=np.array([len(y) for y in ys], np.int32), eouts=out['xs'].detach().numpy(),
                elens=out['xlens'].numpy().astype(np.int32), loss=np.array(float(loss.detach()), np.float64),
                workload=np.array(wname))
    gn = 0.0
    for pre, mod in (("enc.", enc), ("ctc.", ctc)):
        for k, p in mod.named_parameters():
            s = grad_summary(p.grad)
            gn += s["norm"] ** 2
            save["gn." + pre + k] = np.array([s["norm"], s["amax"]], np.float64)
            save["gh." + pre + k] = s["head"]
            save["gs." + pre + k] = s["strided"]
    save["grad_norm"] = np.array(gn ** 0.5, np.float64)
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **save)
    print(tag, "loss %.6f" % float(loss), "eouts", tuple(out['xs'].shape), "elens", out['xlens'].tolist(), "grad norm %.4e" % gn ** 0.5)


def gen_smoke():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from gen_golden_encoder import build_reference
    ctc_mod = importlib.import_module('neural_sp.models.seq2seq.decoders.ctc')
    enc, args, conv_args, kind = build_reference("enc_conformer_small")
    g = np.load(os.path.join(HERE, "enc_conformer_small.npz"))
    V = 40
    gen = torch.Generator().manual_seed(7)
    head = {"output.weight": (torch.rand(V, args["d_model"], generator=gen) * 2 - 1) * 0.1, "output.bias": torch.zeros(V)}
    ctc = ctc_mod.CTC(eos=2, blank=0, enc_n_units=args["d_model"], vocab=V, lsm_prob=0.1)
    ctc.load_state_dict(head, strict=True)
    enc.train(), ctc.train()
    ys = [[5, 6, 7], [8, 9], [10]]
    out, loss = reference_step(enc, ctc, g["xs"], g["xlens"].tolist(), ys)
    sq = {k: float(p.grad.double().pow(2).sum()) for k, p in list(enc.named_parameters()) + [("ctc." + k, p) for k, p in ctc.named_parameters()]}
    gn = sum(sq.values()) ** 0.5
    top = sorted(sq.items(), key=lambda kv: -kv[1])[:3]
    np.savez_compressed(os.path.join(HERE, "bench_smoke.npz"), head_w=head["output.weight"].numpy(), head_b=head["output.bias"].numpy(),
                        loss=np.array(float(loss.detach())), grad_norm=np.array(gn),
                        top=np.array(["%s %.3e" % (k, v ** 0.5) for k, v in top]))
    print("smoke: loss %.6f grad norm %.4e; largest:" % (float(loss), gn), top)


if __name__ == "__main__":
    sys.path.insert(0, HERE)
    gen_smoke()
    gen_bench("bench_c2", "conformer_m_ctc")
    gen_bench("bench_c3", "conformer_l_ctc")
