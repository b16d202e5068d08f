"""Dropout in the training path, CPU part (host wiring): the encoder blocks' dropout call sites, their order, the masked
bias gradients and the mask regeneration in the backward are pinned to the UNMODIFIED reference's autograd.

Method: ours runs first (ops replaced by tests/ops_doubles.py; the dropout double is a numpy restatement of the Philox
counter scheme of csrc/dropout.cu) and every forward dropout call is logged (numel, p, stream id).  The reference then runs
in train() mode with `torch.nn.functional.dropout` replaced by a function that REPLAYS those masks in call order -- the
reference draws its masks in the same order over the same contiguous `[B, T, d]` tensors -- so both sides compute the same
function and outputs / parameter gradients must agree to fp32 rounding.  That the CUDA kernel produces the same Philox
stream as the numpy restatement is checked on the GPU (tests/test_zz_dropout_gpu.py).
Needs /root/reference (build container only): skipped elsewhere."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/neural_sp"), reason="reference tree not available")

BASE = dict(input_dim=80, enc_type='conv_conformer', n_heads=2, kernel_size=7, normalization='layer_norm', n_layers=2,
            n_layers_sub1=0, n_layers_sub2=0, d_model=32, d_ff=64, ffn_bottleneck_dim=0, ffn_activation='swish',
            pe_type='relative', layer_norm_eps=1e-12, last_proj_dim=0, dropout_in=0.0, dropout=0.2, dropout_att=0.0,
            dropout_layer=0.0, subsample="2_1", subsample_type='max_pool', n_stacks=1, n_splices=1, frontend_conv=None,
            task_specific_layer=False, param_init='xavier_uniform', clamp_len=10, lookahead="0_0", chunk_size_left="0",
            chunk_size_current="0", chunk_size_right="0", streaming_type='mask')
CONV = dict(input_dim=80, in_channel=1, channels="32_32", kernel_sizes="(3,3)_(3,3)", strides="(1,1)_(1,1)",
            poolings="(2,2)_(2,2)", dropout=0.0, normalization='', residual=False, bottleneck_dim=32, param_init=0.1)
CASES = {
    "conformer": dict(),
    "conformer_v2_glu": dict(enc_type='conv_conformer_v2', ffn_activation='glu'),
    "uni_conformer": dict(enc_type='conv_uni_conformer', dropout=0.1),
    "transformer_xl": dict(enc_type='conv_transformer', pe_type='relative_xl', ffn_activation='relu', KIND='transformer'),
    "transformer_add": dict(enc_type='conv_transformer', pe_type='add', ffn_activation='gelu', dropout_in=0.15,
                            last_proj_dim=24, KIND='transformer'),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_dropout_training_matches_reference_with_replayed_masks(name, monkeypatch):
    import ops_doubles
    from oracle.ref_import import import_reference
    import_reference()
    from neural_sp_b200 import ops, random as nrandom
    from neural_sp_b200.encoders.conformer import ConformerEncoder
    from neural_sp_b200.encoders.conv import ConvEncoder
    from neural_sp_b200.encoders.transformer import TransformerEncoder
    ops_doubles.install_training(monkeypatch)
    nrandom.manual_seed(1234)
    torch.manual_seed(0)
    ov = dict(CASES[name])
    kind = ov.pop("KIND", "conformer")
    a_ref = dict(BASE)
    a_ref.update(ov)
    a_our = dict(a_ref)
    a_ref['frontend_conv'] = importlib.import_module('neural_sp.models.seq2seq.encoders.conv').ConvEncoder(**CONV)
    a_our['frontend_conv'] = ConvEncoder(**CONV)
    if kind == 'conformer':
        ref = importlib.import_module('neural_sp.models.seq2seq.encoders.conformer').ConformerEncoder(**a_ref)
        ours = ConformerEncoder(**a_our)
    else:
        for a in (a_ref, a_our):
            a.pop("kernel_size"), a.pop("normalization")
        ref = importlib.import_module('neural_sp.models.seq2seq.encoders.transformer').TransformerEncoder(**a_ref)
        ours = TransformerEncoder(**a_our)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.set_precision("fp32")
    ref.train(), ours.train()

    # ---- ours, logging the forward dropout calls ----
    log, rec = [], {"on": True}
    d_drop, d_add = ops_doubles.dropout, ops_doubles.dropout_add

    def drop_logged(x, p, stream_id, scale=1.0, out_dtype=None, inplace=False):
        if rec["on"]:
            log.append((x.numel(), float(p), int(stream_id)))
        return d_drop(x, p, stream_id, scale, out_dtype, inplace)

    def add_logged(t, res, p, alpha, stream_id, out=None):
        if rec["on"]:
            log.append((t.numel(), float(p), int(stream_id)))
        rec_on, rec["on"] = rec["on"], False            # dropout_add's double calls dropout itself
        y = d_add(t, res, p, alpha, stream_id, out)
        rec["on"] = rec_on
        return y

    monkeypatch.setattr(ops, "dropout", drop_logged)
    monkeypatch.setattr(ops, "dropout_add", add_logged)
    monkeypatch.setattr(ops_doubles, "dropout", drop_logged)      # resolved at call time inside dropout_add's double

    rng = np.random.RandomState(3)
    xs = torch.from_numpy(rng.randn(2, 60, 80).astype(np.float32))
    xs[1, 50:] = 0
    xlens = torch.IntTensor([60, 50])
    out = ours(xs.clone(), xlens.clone(), task='all')['ys']
    ys = out['xs']
    rec["on"] = False
    n_layers = a_our['n_layers']
    per_block = {"conformer": 6, "transformer": 3}[kind]
    assert len(log) >= n_layers * per_block, (len(log), log[:4])
    w = torch.from_numpy(np.random.RandomState(5).randn(*ys.shape).astype(np.float32))
    for b, n in enumerate(out['xlens'].tolist()):
        w[b, n:] = 0
    (ys * w).sum().backward()

    # ---- reference with the same masks replayed in call order ----
    seed, off = int(nrandom.state("cpu")[0]), int(nrandom.state("cpu")[1])
    it = iter(log)

    def replay(input, p=0.5, training=True, inplace=False):
        if not training or p == 0:
            return input
        numel, p_rec, sid = next(it)
        assert numel == input.numel() and abs(p_rec - p) < 1e-7, ((numel, p_rec), (input.numel(), p))
        keep = ops_doubles.philox_keep(numel, float(torch.tensor(p, dtype=torch.float32)), seed, off, sid).reshape(input.shape)
        s = torch.tensor(1.0) / (1 - torch.tensor(p, dtype=torch.float32))
        return torch.where(keep, input * s, torch.zeros(()))

    monkeypatch.setattr(torch.nn.functional, "dropout", replay)
    r_out = ref(xs.clone(), xlens.clone(), task='all')['ys']
    assert next(it, None) is None, "ours made more dropout calls than the reference"
    assert torch.equal(r_out['xlens'], out['xlens'])
    scale = float(r_out['xs'].abs().max())
    assert float((r_out['xs'] - ys).abs().max()) <= 1e-4 * scale, float((r_out['xs'] - ys).abs().max())
    (r_out['xs'] * w).sum().backward()
    ref_g = {k: p.grad for k, p in ref.named_parameters()}
    gmax = max(float(g.abs().max()) for g in ref_g.values() if g is not None)
    bad = []
    for k, p in ours.named_parameters():
        g = ref_g[k]
        if g is None:
            continue
        assert p.grad is not None, k
        e = float((p.grad - g).abs().max() / max(float(g.abs().max()), 1e-3 * gmax))
        if not e <= 1e-3:
            bad.append((k, e))
    assert not bad, (name, bad[:8], len(bad))


def test_ctc_head_dropout_matches_reference_with_replayed_masks(monkeypatch):
    """CTC head with `fc_list` (Linear -> Dropout -> Linear, ctc.py:82-89) + label smoothing in train mode: loss, gradient
    w.r.t. the encoder output and the head's parameter gradients against the unmodified reference `CTC` with our masks."""
    import ops_doubles
    from oracle.ref_import import import_reference
    import_reference()
    from neural_sp.models.seq2seq.decoders.ctc import CTC as RefCTC
    from neural_sp_b200 import ops, random as nrandom
    from neural_sp_b200.decoders.ctc import CTC
    ops_doubles.install_training(monkeypatch)
    nrandom.manual_seed(5)
    torch.manual_seed(0)
    kw = dict(eos=2, blank=0, enc_n_units=24, vocab=30, dropout=0.25, lsm_prob=0.1, fc_list="16_12")
    ref, ours = RefCTC(**kw).train(), CTC(**kw)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.set_precision("fp32")
    ours.train()
    log = []
    d_drop = ops_doubles.dropout

    def drop_logged(x, p, stream_id, scale=1.0, out_dtype=None, inplace=False):
        log.append((x.numel(), float(p), int(stream_id)))
        return d_drop(x, p, stream_id, scale, out_dtype, inplace)

    monkeypatch.setattr(ops, "dropout", drop_logged)
    e0 = torch.randn(3, 20, 24)
    elens = torch.IntTensor([20, 17, 12])
    ys = [[5, 6, 7, 8], [9, 10], [4, 4, 11]]
    e_our = e0.clone().requires_grad_(True)
    loss_o, _ = ours(e_our, elens.clone(), ys)
    n_fwd = len(log)
    assert n_fwd == 2
    loss_o.backward()
    seed, off = int(nrandom.state("cpu")[0]), int(nrandom.state("cpu")[1])
    it = iter(log[:n_fwd])

    def replay(input, p=0.5, training=True, inplace=False):
        if not training or p == 0:
            return input
        numel, p_rec, sid = next(it)
        assert numel == input.numel()
        keep = ops_doubles.philox_keep(numel, float(torch.tensor(p, dtype=torch.float32)), seed, off, sid).reshape(input.shape)
        return torch.where(keep, input / (1 - torch.tensor(p, dtype=torch.float32)), torch.zeros(()))

    monkeypatch.setattr(torch.nn.functional, "dropout", replay)
    e_ref = e0.clone().requires_grad_(True)
    loss_r, _ = ref(e_ref, elens.clone(), ys)
    loss_r.backward()
    assert abs(float(loss_o.detach()) - float(loss_r.detach())) <= 1e-4 * abs(float(loss_r.detach()))
    assert float((e_our.grad - e_ref.grad).abs().max()) <= 1e-4 * float(e_ref.grad.abs().max())
    rg = dict(ref.named_parameters())
    for k, p in ours.named_parameters():
        err = float((p.grad - rg[k].grad).abs().max() / rg[k].grad.abs().max().clamp_min(1e-12))
        assert err <= 2e-4, (k, err)
