"""CNN front-end (ConvEncoder / Conv2dBlock, reference conv.py:18-396) host logic on CPU against the UNMODIFIED reference:
the reference's own test matrix (test/encoders/test_conv_encoder.py: poolings incl. (2,1) / (1,1), three blocks, BatchNorm2d,
LayerNorm2D, residual, bottleneck) plus strided blocks ("ConvSubsample", used by its streaming tests), in eval mode with
non-trivial BatchNorm running statistics.  Ops replaced by their torch restatements (tests/ops_doubles.py).
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


def make_args(**kw):
    a = dict(input_dim=80, in_channel=1, channels="32_32_32", kernel_sizes="(3,3)_(3,3)_(3,3)", strides="(1,1)_(1,1)_(1,1)",
             poolings="(2,2)_(2,2)_(2,2)", dropout=0.1, normalization='', residual=False, bottleneck_dim=0, param_init=0.1)
    a.update(kw)
    return a


TWO = dict(channels="32_32", kernel_sizes="(3,3)_(3,3)", strides="(1,1)_(1,1)")
CASES = [
    dict(TWO, poolings="(2,2)_(2,2)"), dict(TWO, poolings="(2,2)_(2,1)"), dict(TWO, poolings="(1,1)_(1,1)"),
    dict(poolings="(2,2)_(2,2)_(2,2)"), dict(poolings="(2,2)_(2,2)_(2,1)"), dict(poolings="(2,2)_(2,1)_(2,1)"),
    dict(poolings="(2,2)_(1,1)_(1,1)"), dict(poolings="(2,1)_(1,1)_(1,1)"), dict(poolings="(1,1)_(1,1)_(1,1)"),
    dict(normalization='batch_norm'), dict(normalization='layer_norm'), dict(residual=True), dict(bottleneck_dim=8),
    dict(normalization='batch_norm', residual=True, bottleneck_dim=16),
    # strided convolutions instead of pooling
    dict(channels="32", kernel_sizes="(3,3)", strides="(2,2)", poolings="(1,1)"),
    dict(TWO, strides="(2,2)_(2,2)", poolings="(1,1)_(1,1)"),
    dict(TWO, strides="(1,1)_(2,2)", poolings="(1,1)_(1,1)", normalization='layer_norm'),
    dict(strides="(2,2)_(2,2)_(2,2)", poolings="(1,1)_(1,1)_(1,1)", bottleneck_dim=24),
]


@pytest.mark.parametrize("ov", CASES)
@pytest.mark.parametrize("look", [(False, False), (True, True)])
def test_conv_encoder_matches_reference(ov, look, monkeypatch):
    import ops_doubles
    from oracle.ref_import import import_reference
    import_reference()
    from neural_sp_b200.encoders.conv import ConvEncoder
    ops_doubles.install(monkeypatch)
    torch.manual_seed(0)
    args = make_args(**ov)
    ref = importlib.import_module('neural_sp.models.seq2seq.encoders.conv').ConvEncoder(**args)
    g = torch.Generator().manual_seed(1)
    for m in ref.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    ours = ConvEncoder(**args)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.set_precision("fp32")
    ref.eval(), ours.eval()
    assert ours.output_dim == ref.output_dim and ours.subsampling_factor == ref.subsampling_factor
    assert ours.context_size == ref.context_size
    rng = np.random.RandomState(0)
    for xmax in (40, 45):
        xs = torch.from_numpy(rng.randn(4, xmax, 80).astype(np.float32))
        xlens = torch.IntTensor([xmax - 3 * i for i in range(4)])
        for b, n in enumerate(xlens.tolist()):
            xs[b, n:] = 0
        with torch.no_grad():
            r_xs, r_lens = ref(xs.clone(), xlens.clone(), lookback=look[0], lookahead=look[1])
        o_xs, o_lens = ours(xs.clone(), xlens.clone(), lookback=look[0], lookahead=look[1])
        assert torch.equal(r_lens, o_lens), (r_lens, o_lens)
        assert r_xs.shape == o_xs.shape, (r_xs.shape, o_xs.shape)
        assert torch.allclose(r_xs, o_xs, atol=1e-4), float((r_xs - o_xs).abs().max())


# Round 1 left cases 2 and 7 (no bridge, un-pooled first block) as an open 0.2-0.6 % gradient deviation.  Bisected in round 2:
# ONE ReLU mask bit.  With the seed-0 input the reference's pre-activation layers.1.conv1[0, 31, 5, 53] is -6.7e-8 while the
# restatement's summation order gives +1.3e-7; the upstream gradient there is -0.70, so that single element changes the 288
# weight-gradient entries of output channel 31 by 0.3-0.5 % of the tensor's maximum (and everything upstream of it).  Neither
# side is wrong -- ReLU'(0+-eps) is decided by rounding.  These two cases (four times the positions of the pooled stacks, hence
# the collision) use another input seed; every case keeps the "<= 2 entries above 3e-3" allowance for the same effect.
_SEED = {2: 1, 7: 1}


@pytest.mark.parametrize("ov", CASES)
def test_conv_encoder_training_matches_reference(ov, monkeypatch):
    """The same matrix in train() mode (dropout 0): the REAL training node (autograd._FrontendFn: its forward and its
    hand-written backward chain, ops replaced by their restatements) against torch autograd over the unmodified reference --
    output, every parameter gradient, BatchNorm's running statistics after the step."""
    import ops_doubles
    from oracle.ref_import import import_reference
    import_reference()
    from neural_sp_b200 import autograd as ag
    from neural_sp_b200.encoders.conv import ConvEncoder
    real_node = ag.frontend_forward
    ops_doubles.install_training(monkeypatch)
    monkeypatch.setattr(ag, "frontend_forward", real_node)
    torch.manual_seed(0)
    args = make_args(**ov)
    args['dropout'] = 0.0
    ref = importlib.import_module('neural_sp.models.seq2seq.encoders.conv').ConvEncoder(**args).train()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for k, p in ref.named_parameters():
            if ".norm" in k:                                  # (1-D parameters are initialised to 0: gamma = 0 would kill the net)
                p.add_(0.5 + 0.3 * torch.randn(p.shape, generator=g))
    ours = ConvEncoder(**args)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.set_precision("fp32")
    ours.train()
    rng = np.random.RandomState(_SEED.get(CASES.index(ov), 0))
    xs = torch.from_numpy(rng.randn(4, 45, 80).astype(np.float32))
    xlens = torch.IntTensor([45 - 3 * i for i in range(4)])
    for b, n in enumerate(xlens.tolist()):
        xs[b, n:] = 0
    r_xs, r_lens = ref(xs.clone(), xlens.clone())
    o_xs = ag.frontend_forward(ours, xs.clone(), 1.0, "fp32")
    assert torch.equal(r_lens, ours.output_lens(xlens)) and r_xs.shape == o_xs.shape
    assert float((r_xs - o_xs).abs().max()) <= 1e-4 * float(r_xs.abs().max())
    w = torch.from_numpy(rng.randn(*r_xs.shape).astype(np.float32))
    (r_xs * w).sum().backward()
    (o_xs * w).sum().backward()
    rg = dict(ref.named_parameters())
    gmax = max(float(p.grad.abs().max()) for p in rg.values())
    for k, p in ours.named_parameters():
        d = (p.grad - rg[k].grad).abs() / max(float(rg[k].grad.abs().max()), 1e-2 * gmax)
        # (<= 2 entries per tensor may sit on a ReLU mask bit that rounding flips: see tests/test_frontend_node_cpu.py)
        assert int((d > 3e-3).sum()) <= 2 and float(d.max()) <= 5e-2, (k, float(d.max()), int((d > 3e-3).sum()))
    for (k, b), (_, rb) in zip(ours.named_buffers(), ref.named_buffers()):
        assert torch.allclose(b.float(), rb.float(), rtol=1e-4, atol=1e-5), k
