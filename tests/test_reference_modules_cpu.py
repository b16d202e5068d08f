"""The reference's module-level test matrices that fall on the hot path (test/modules/test_conformer_convolution.py and
test_pointwise_feed_forward.py ::test_forward) through neural_sp_b200's modules called the reference's way (`module(xs)`), eval
mode, against the unmodified reference modules with identical weights.  CPU: ops replaced by their torch restatements.
Kernel sizes up to 65, causal variant, the three normalisations, every FFN activation, the low-rank form.
Needs /root/reference (build container only): skipped elsewhere."""
import importlib
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/neural_sp"), reason="reference tree not available")


def _pair(ref_mod, ref_cls, ours_cls, kw, monkeypatch):
    import ops_doubles
    from oracle.ref_import import import_reference
    import_reference()
    ops_doubles.install(monkeypatch)
    torch.manual_seed(0)
    ref = getattr(importlib.import_module(ref_mod), ref_cls)(**kw).eval()
    ours = ours_cls(**kw)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.precision = "fp32"
    for m in ours.modules():
        m.precision = "fp32"
    return ref, ours.eval()


@pytest.mark.parametrize("ov", [{'kernel_size': 3}, {'kernel_size': 7}, {'kernel_size': 17}, {'kernel_size': 31}, {'kernel_size': 33},
                                {'kernel_size': 65}, {'param_init': 'xavier_uniform'}, {'param_init': 'lecun'},
                                {'kernel_size': 7, 'causal': True}, {'normalization': 'group_norm'}, {'normalization': 'layer_norm'}])
def test_conformer_convolution_matrix(ov, monkeypatch):
    from neural_sp_b200.modules.conformer_convolution import ConformerConvBlock
    kw = dict(d_model=256, kernel_size=3, param_init='', causal=False, normalization='batch_norm')
    kw.update(ov)
    ref, ours = _pair('neural_sp.models.modules.conformer_convolution', 'ConformerConvBlock', ConformerConvBlock, kw, monkeypatch)
    g = torch.Generator().manual_seed(1)
    if kw['normalization'] == 'batch_norm':
        for m in (ref, ours):
            m.norm.running_mean.copy_(torch.randn(256, generator=torch.Generator().manual_seed(1)) * 0.1)
            m.norm.running_var.copy_(torch.rand(256, generator=torch.Generator().manual_seed(2)) + 0.5)
    for xmax in (40, 45):
        xs = torch.randn(4, xmax, 256, generator=g)
        with torch.no_grad():
            r = ref(xs.clone())
        o = ours(xs.clone())
        assert o.shape == r.shape == (4, xmax, 256)
        assert float((o - r).abs().max()) <= 1e-4 * float(r.abs().max())


@pytest.mark.parametrize("ov", [{'activation': 'relu'}, {'activation': 'gelu'}, {'activation': 'gelu_accurate'}, {'activation': 'glu'},
                                {'activation': 'swish'}, {'param_init': 'xavier_uniform'}, {'bottleneck_dim': 16}])
def test_positionwise_feed_forward_matrix(ov, monkeypatch):
    from neural_sp_b200.modules.positionwise_feed_forward import PositionwiseFeedForward
    kw = dict(d_model=32, d_ff=128, dropout=0.1, activation='relu', param_init='', bottleneck_dim=0)
    kw.update(ov)
    ref, ours = _pair('neural_sp.models.modules.positionwise_feed_forward', 'PositionwiseFeedForward', PositionwiseFeedForward, kw,
                      monkeypatch)
    xs = torch.randn(4, 40, 32, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        r = ref(xs.clone())
    o = ours(xs.clone())
    assert o.shape == r.shape and float((o - r).abs().max()) <= 1e-4 * float(r.abs().max())
