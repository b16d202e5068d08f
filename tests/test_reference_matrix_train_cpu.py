"""The reference's own encoder test matrices in TRAINING mode (train(), dropouts set to 0 so that both sides compute the same
function): neural_sp_b200's autograd nodes (ops replaced by their torch restatements, tests/ops_doubles.py) against torch
autograd over the UNMODIFIED reference with identical weights -- outputs (incl. sub-task outputs) and every parameter
gradient.  Configurations whose training path is not on the B200 path raise NotImplementedError and are reported as skips
(the inference parity of ALL configurations is tests/test_reference_matrix_cpu.py); the skip reasons are the honest list of
training gaps (today: none; one configuration is skipped because the reference's own gradients are not finite).
Needs /root/reference (build container only): skipped elsewhere."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_reference_matrix_cpu import FAMILIES, _matrix  # noqa: E402

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/neural_sp"), reason="reference tree not available")


@pytest.mark.parametrize("family, ov, ov_conv", _matrix())
def test_reference_test_matrix_training_parity(family, ov, ov_conv, monkeypatch):
    import ops_doubles
    from neural_sp_b200.encoders.conformer import ConformerEncoder
    from neural_sp_b200.encoders.conv import ConvEncoder
    from neural_sp_b200.encoders.rnn import RNNEncoder
    from neural_sp_b200.encoders.transformer import TransformerEncoder
    ops_doubles.install_training(monkeypatch)
    tmod, rmod, cls = FAMILIES[family]
    tm = importlib.import_module(tmod)
    ours_cls = {"conformer": ConformerEncoder, "transformer": TransformerEncoder, "rnn": RNNEncoder}[family]
    args = tm.make_args(**ov)
    for k in ('dropout', 'dropout_att', 'dropout_in', 'dropout_layer', 'rsp_prob'):
        if k in args:
            args[k] = 0.0
    a_ref, a_our = dict(args), dict(args)
    torch.manual_seed(0)
    if 'conv' in args['enc_type']:
        c = tm.make_args_conv(**ov_conv)
        c['dropout'] = 0.0
        if family != 'rnn':
            c['bottleneck_dim'] = args['d_model']
        a_ref['frontend_conv'] = importlib.import_module('neural_sp.models.seq2seq.encoders.conv').ConvEncoder(**c)
        a_our['frontend_conv'] = ConvEncoder(**c)
    ref = getattr(importlib.import_module(rmod), cls)(**a_ref).train()
    ours = ours_cls(**a_our)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.set_precision('fp32')
    ours.train()
    rng = np.random.RandomState(0)
    lc = str(args.get('chunk_size_current', '0')) not in ('0',)
    xmax = 90 if (lc or family == 'rnn') else 45
    xs = torch.from_numpy(rng.randn(4, xmax, args['input_dim']).astype(np.float32))
    xlens = torch.IntTensor([xmax - i * ref.subsampling_factor for i in range(4)])
    for b, n in enumerate(xlens.tolist()):
        xs[b, n:] = 0
    try:
        o = ours(xs.clone(), xlens.clone(), task='all')
    except NotImplementedError as e:
        pytest.skip("no training path on the B200 path: %s" % str(e)[:90])
    r = ref(xs.clone(), xlens.clone(), task='all')
    lr = lo = 0
    for k in ('ys', 'ys_sub1', 'ys_sub2'):
        if r[k]['xs'] is None:
            continue
        assert r[k]['xs'].shape == o[k]['xs'].shape and torch.equal(torch.as_tensor(r[k]['xlens']), torch.as_tensor(o[k]['xlens']))
        err = float((r[k]['xs'] - o[k]['xs']).abs().max() / r[k]['xs'].abs().max().clamp_min(1e-6))
        assert err <= 1e-4, (k, err)
        w = torch.from_numpy(np.random.RandomState(7).randn(*r[k]['xs'].shape).astype(np.float32))
        for b, n in enumerate(torch.as_tensor(r[k]['xlens']).tolist()):
            w[b, n:] = 0
        lr, lo = lr + (r[k]['xs'] * w).sum(), lo + (o[k]['xs'] * w).sum()
    lr.backward()
    lo.backward()
    rg = dict(ref.named_parameters())
    nonfinite = sorted(k for k, p in rg.items() if p.grad is not None and not bool(torch.isfinite(p.grad).all()))
    if nonfinite:
        # the REFERENCE's own autograd overflows here (BatchNorm2d front-end + the eps = 1e-12 LayerNorms of the Conformer
        # blocks at these toy sizes: NaN / 1e13 gradients); forward parity held above, and the overflow is reproduced:
        ours_nonfinite = sorted(k for k, p in ours.named_parameters() if p.grad is not None and not bool(torch.isfinite(p.grad).all()))
        assert ours_nonfinite == nonfinite
        pytest.skip("the reference's gradients are not finite for this configuration (%d tensors; same set on both sides)" % len(nonfinite))
    gmax = max(float(p.grad.abs().max()) for p in rg.values() if p.grad is not None)
    bad = []
    for k, p in ours.named_parameters():
        g = rg[k].grad
        if g is None:
            continue
        assert p.grad is not None, k
        e = float((p.grad - g).abs().max() / max(float(g.abs().max()), 1e-3 * gmax))
        if not e <= 1e-3:
            bad.append((k, e))
    assert not bad, (bad[:8], len(bad))
