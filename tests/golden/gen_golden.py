#!/usr/bin/env python3
"""Generate golden input/output vectors by RUNNING THE UNMODIFIED REFERENCE on CPU.

Run in the build container (the reference lives at /root/reference and cannot travel to the
GPU box):   python tests/golden/gen_golden.py [ctc] [encoder] [rnnt]

The reference's own tests hold no value-level vectors for this path, so
these fixtures are how the oracle (oracle/*.py) and the CUDA path are pinned to the reference.
Every fixture stores the seeded inputs, the parameters and the reference's outputs.
"""
import os
import zlib
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_import import import_reference  # noqa: E402

import_reference()


def _rand_labels(rng, B, ylens, V, force_repeats=False):
    ys = []
    for b in range(B):
        y = rng.integers(4, V, size=ylens[b]).tolist()  # 0-3 reserved (speech2text.py:66-69)
        if force_repeats and len(y) >= 2:
            y[1] = y[0]
        ys.append(y)
    return ys


def gen_ctc():
    from neural_sp.models.seq2seq.decoders.ctc import CTC
    from neural_sp.models.criterion import kldiv_lsm_ctc

    cases = {
        # name: (B, T, V, elens, ylens, lsm, repeats, scale)
        "c1_blstm_shape": (2, 200, 32, [200, 180], [30, 22], 0.0, False, 1.0),      # BASELINE configs[0]
        "las_test_shape": (4, 40, 10, [40, 38, 36, 34], [4, 5, 3, 7], 0.1, True, 1.0),  # test_las_decoder.py:115
        "ragged_repeats": (5, 61, 50, [61, 33, 7, 1, 50], [20, 11, 3, 1, 0], 0.1, True, 2.0),
        "infeasible": (3, 12, 20, [12, 5, 12], [4, 6, 12], 0.0, True, 1.0),         # zero_infinity
        "bpe1k_small": (4, 125, 1000, [125, 120, 99, 64], [56, 50, 41, 20], 0.1, False, 3.0),
    }
    for name, (B, T, V, elens, ylens, lsm, rep, scale) in cases.items():
        rng = np.random.default_rng(zlib.crc32(name.encode()))
        torch.manual_seed(0)
        logits = (torch.randn(B, T, V) * scale).requires_grad_(True)
        ys = _rand_labels(rng, B, ylens, V, force_repeats=rep)
        if name == "infeasible":
            ys[2] = [5] * 12          # needs 12 + 11 frames > 12
        ctc = CTC(eos=2, blank=0, enc_n_units=8, vocab=V, lsm_prob=lsm)
        ctc.train()
        elens_t = torch.IntTensor(elens)
        ylens_t = torch.IntTensor([len(y) for y in ys])
        ys_ctc = torch.cat([torch.IntTensor(y) for y in ys if len(y) > 0] or [torch.IntTensor([])])
        # the reference's op boundary: loss_fn(logits [T,B,V], ...) ctc.py:139-150
        loss_ctc = ctc.loss_fn(logits.transpose(1, 0), ys_ctc, elens_t, ylens_t)
        loss = loss_ctc
        kl = torch.zeros(())
        if lsm > 0:
            kl = kldiv_lsm_ctc(logits, elens_t)
            loss = loss_ctc * (1 - lsm) + kl * lsm
        loss.backward()
        out = dict(logits=logits.detach().numpy(), elens=np.array(elens, np.int32),
                   ylens=ylens_t.numpy(), ys_cat=ys_ctc.numpy().astype(np.int32),
                   lsm=np.float32(lsm), loss_ctc=loss_ctc.detach().numpy(), kl=kl.detach().numpy(),
                   loss=loss.detach().numpy(), grad=logits.grad.numpy())
        # per-utterance nll (reduction none) for diagnostics
        nll = torch.nn.functional.ctc_loss(logits.detach().transpose(1, 0).log_softmax(2), ys_ctc, elens_t, ylens_t,
                                           reduction="none", zero_infinity=True)
        out["nll"] = nll.numpy()
        if name != "infeasible" and min(len(y) for y in ys) > 0:
            with torch.no_grad():
                trig = ctc.forced_aligner(logits.detach().clone(), elens_t, ys, ylens_t)
            out["trigger_points"] = trig.numpy().astype(np.int32)
        np.savez_compressed(os.path.join(HERE, "ctc_%s.npz" % name), **out)
        print("ctc", name, float(loss), out["grad"].shape, "trig" in "".join(out.keys()))

    # CTC.forward through the Linear head (fc_list) : ctc.py:105-137
    torch.manual_seed(1)
    B, T, D, V = 3, 50, 24, 40
    ctc = CTC(eos=2, blank=0, enc_n_units=D, vocab=V, lsm_prob=0.1, fc_list="16")
    ctc.train()
    eouts = torch.randn(B, T, D, requires_grad=True)
    rng = np.random.default_rng(7)
    ys = _rand_labels(rng, B, [10, 8, 5], V)
    elens_t = torch.IntTensor([50, 44, 30])
    loss, trig = ctc(eouts, elens_t, ys, forced_align=True)
    loss.backward()
    sd = {"sd." + k: v.detach().numpy() for k, v in ctc.state_dict().items()}
    gsd = {"gsd." + k: p.grad.numpy() for k, p in ctc.named_parameters()}
    np.savez_compressed(os.path.join(HERE, "ctc_head_forward.npz"), eouts=eouts.detach().numpy(),
                        elens=elens_t.numpy(), ylens=np.array([len(y) for y in ys], np.int32),
                        ys_cat=np.concatenate(ys).astype(np.int32), loss=loss.detach().numpy(),
                        grad_eouts=eouts.grad.numpy(), trigger_points=trig.numpy().astype(np.int32), **sd, **gsd)
    print("ctc head forward", float(loss))


if __name__ == "__main__":
    what = sys.argv[1:] or ["ctc", "encoder", "rnn", "rnnt"]
    if "ctc" in what:
        gen_ctc()
    if "encoder" in what:
        from gen_golden_encoder import gen_encoder
        gen_encoder()
    if "encoder_grads" in what:
        from gen_golden_encoder import gen_encoder_grads
        gen_encoder_grads()
    if "rnn_grads" in what:
        from gen_golden_encoder import gen_rnn_grads
        gen_rnn_grads()
    if "rnn" in what:
        from gen_golden_encoder import gen_rnn_encoder
        gen_rnn_encoder()
    if "rnnt" in what:
        from gen_golden_encoder import gen_rnnt
        gen_rnnt()
