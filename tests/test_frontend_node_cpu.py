"""The CNN front-end's training node (neural_sp_b200/autograd.py _FrontendFn: the hand-written backward chain -- ReLU / pooling
masks, input-gradient convolutions with flipped taps, weight gradients, the bridge's permuted columns, the strided blocks'
scatter back onto the stride-1 grid) on CPU: the REAL node with its ops replaced by torch restatements, against torch autograd
through the plain conv stack.  Runs wherever the repository does."""
import numpy as np
import pytest
import torch


def _cfg(**kw):
    a = dict(input_dim=80, in_channel=1, channels="32_32", kernel_sizes="(3,3)_(3,3)", strides="(1,1)_(1,1)",
             poolings="(2,2)_(2,2)", dropout=0.0, normalization='', residual=False, bottleneck_dim=0, param_init=0.1)
    a.update(kw)
    return a


@pytest.mark.parametrize("ov", [
    dict(), dict(poolings="(1,1)_(2,2)", bottleneck_dim=24), dict(poolings="(2,2)_(2,1)"), dict(poolings="(1,1)_(1,1)", bottleneck_dim=16),
    dict(strides="(2,2)_(2,2)", poolings="(1,1)_(1,1)", bottleneck_dim=24), dict(strides="(1,1)_(2,2)", poolings="(1,1)_(1,1)"),
    dict(strides="(2,2)_(1,1)", poolings="(1,1)_(2,2)", bottleneck_dim=8),
    dict(channels="32", kernel_sizes="(3,3)", strides="(2,2)", poolings="(1,1)"),
    dict(input_dim=240, in_channel=3),
])
def test_frontend_node_matches_torch_autograd(ov, monkeypatch):
    import ops_doubles
    from neural_sp_b200 import autograd as ag
    from neural_sp_b200.encoders.conv import ConvEncoder
    real_node = ag.frontend_forward                      # keep the real node; install_training replaces the module attribute
    ops_doubles.install_training(monkeypatch)
    torch.manual_seed(0)
    enc = ConvEncoder(**_cfg(**ov)).train()
    enc.set_precision("fp32")
    rng = np.random.RandomState(1)
    xs = torch.from_numpy(rng.randn(3, 37, enc.in_channel * enc.input_freq).astype(np.float32))
    y_ref = ops_doubles.frontend_forward(enc, xs, 1.7, "fp32")
    w = torch.from_numpy(rng.randn(*y_ref.shape).astype(np.float32))
    (y_ref * w).sum().backward()
    ref = {k: p.grad.clone() for k, p in enc.named_parameters()}
    enc.zero_grad()
    y = real_node(enc, xs, 1.7, "fp32")
    assert y.shape == y_ref.shape
    assert float((y - y_ref).abs().max()) <= 1e-5 * float(y_ref.abs().max())
    (y * w).sum().backward()
    for k, p in enc.named_parameters():
        err = float((p.grad - ref[k]).abs().max() / ref[k].abs().max().clamp_min(1e-12))
        assert err <= 2e-4, (k, err)
    lens = enc.output_lens(torch.IntTensor([37, 30, 11]))
    assert lens.tolist()[0] == y.shape[1]


def _ln2d_stack(enc, xs, out_scale):
    """torch restatement of the LayerNorm2D front-end (reference conv.py:362-421): conv -> LN over a frame's [C, F] -> ReLU."""
    F_ = torch.nn.functional
    B, T, Fd = xs.shape
    x = xs.view(B, T, enc.in_channel, Fd // enc.in_channel).transpose(1, 2)

    def ln(norm, t):                                     # [B, C, T, F] -> LayerNorm([C, F]) per frame
        return norm.norm(t.transpose(1, 2)).transpose(1, 2)
    for blk in enc.layers:
        x = torch.relu(ln(blk.norm1, F_.conv2d(x, blk.conv1.weight, blk.conv1.bias, padding=1)))
        x = torch.relu(ln(blk.norm2, F_.conv2d(x, blk.conv2.weight, blk.conv2.bias, padding=1, stride=tuple(blk.stride))))
        if blk.pool is not None:
            x = F_.max_pool2d(x, blk.pooling, blk.pooling, ceil_mode=True)
    B, C, T, Fq = x.shape
    x = x.transpose(1, 2).reshape(B, T, C * Fq)
    if enc.bridge is not None:
        x = F_.linear(x, enc.bridge.weight, enc.bridge.bias)
    return x * out_scale


@pytest.mark.parametrize("ov", [
    dict(), dict(poolings="(1,1)_(2,2)", bottleneck_dim=24), dict(strides="(1,1)_(2,2)", poolings="(1,1)_(1,1)"),
    dict(strides="(2,2)_(1,1)", poolings="(1,1)_(2,2)", bottleneck_dim=8), dict(input_dim=120, in_channel=3, channels="16_32"),
])
def test_layernorm2d_frontend_node_matches_torch_autograd(ov, monkeypatch):
    """LayerNorm2D blocks in training: LN forward / backward kernels between the convolutions, affine parameters permuted
    between the parameter's [C, F] layout and the channels-last frame's (f, c) order."""
    import ops_doubles
    from neural_sp_b200 import autograd as ag
    from neural_sp_b200.encoders.conv import ConvEncoder
    real_node = ag.frontend_forward
    ops_doubles.install_training(monkeypatch)
    torch.manual_seed(0)
    enc = ConvEncoder(**_cfg(normalization='layer_norm', **ov)).train()
    enc.set_precision("fp32")
    with torch.no_grad():                                # non-trivial affine parameters
        for blk in enc.layers:
            for n in (blk.norm1, blk.norm2):
                n.norm.weight.add_(0.3 * torch.randn_like(n.norm.weight))
                n.norm.bias.add_(0.3 * torch.randn_like(n.norm.bias))
    rng = np.random.RandomState(1)
    xs = torch.from_numpy(rng.randn(3, 37, enc.in_channel * enc.input_freq).astype(np.float32))
    y_ref = _ln2d_stack(enc, xs, 1.7)
    w = torch.from_numpy(rng.randn(*y_ref.shape).astype(np.float32))
    (y_ref * w).sum().backward()
    ref = {k: p.grad.clone() for k, p in enc.named_parameters()}
    enc.zero_grad()
    y = real_node(enc, xs, 1.7, "fp32")
    assert y.shape == y_ref.shape
    assert float((y - y_ref).abs().max()) <= 1e-5 * float(y_ref.abs().max())
    (y * w).sum().backward()
    for k, p in enc.named_parameters():
        d = (p.grad - ref[k]).abs() / ref[k].abs().max().clamp_min(1e-12)
        # a normalised value within rounding of 0 may pass the ReLU in one implementation and not in the other: that flips ONE
        # mask bit and shows up in the one (f, c) entry of the LayerNorm gradients it feeds -- tolerated, at most 2 per tensor
        assert int((d > 1e-3).sum()) <= 2 and float(d.max()) <= 5e-2, (k, float(d.max()), int((d > 1e-3).sum()))


def _general_stack(enc, xs, out_scale):
    """torch restatement of Conv2dBlock stacks with any normalisation and the ResNet-style skip connection (conv.py:360-394)."""
    F_ = torch.nn.functional
    B, T, Fd = xs.shape
    x = xs.view(B, T, enc.in_channel, Fd // enc.in_channel).transpose(1, 2)

    def norm(m, t):
        if m is None:
            return t
        return m(t) if isinstance(m, torch.nn.BatchNorm2d) else m.norm(t.transpose(1, 2)).transpose(1, 2)
    for blk in enc.layers:
        res = x
        x = torch.relu(norm(blk.norm1, F_.conv2d(x, blk.conv1.weight, blk.conv1.bias, padding=1)))
        x = norm(blk.norm2, F_.conv2d(x, blk.conv2.weight, blk.conv2.bias, padding=1, stride=tuple(blk.stride)))
        if blk.residual and x.shape == res.shape:
            x = x + res
        x = torch.relu(x)
        if blk.pool is not None:
            x = F_.max_pool2d(x, blk.pooling, blk.pooling, ceil_mode=True)
    B, C, T, Fq = x.shape
    x = x.transpose(1, 2).reshape(B, T, C * Fq)
    if enc.bridge is not None:
        x = F_.linear(x, enc.bridge.weight, enc.bridge.bias)
    return x * out_scale


@pytest.mark.parametrize("normalization", ['', 'layer_norm', 'batch_norm'])
@pytest.mark.parametrize("ov", [
    dict(channels="32_32_32", kernel_sizes="(3,3)_(3,3)_(3,3)", strides="(1,1)_(1,1)_(1,1)", poolings="(2,2)_(1,1)_(2,2)", bottleneck_dim=24),
    dict(channels="16_16", poolings="(1,1)_(2,2)"),
    dict(channels="32_32_32", kernel_sizes="(3,3)_(3,3)_(3,3)", strides="(1,1)_(2,2)_(1,1)", poolings="(1,1)_(1,1)_(1,1)"),
])
def test_residual_frontend_node_matches_torch_autograd(ov, normalization, monkeypatch):
    """ResNet-style skip connections (live for blocks that keep the activation's shape) with every normalisation kind."""
    import copy
    import ops_doubles
    from neural_sp_b200 import autograd as ag
    from neural_sp_b200.encoders.conv import ConvEncoder
    real_node = ag.frontend_forward
    ops_doubles.install_training(monkeypatch)
    torch.manual_seed(0)
    enc = ConvEncoder(**_cfg(normalization=normalization, residual=True, **ov)).train()
    enc.set_precision("fp32")
    assert any(blk.residual_active for blk in enc.layers)
    with torch.no_grad():               # (the reference's initialiser zeroes every 1-D parameter, BatchNorm's gamma included)
        for k, p in enc.named_parameters():
            if ".norm" in k:
                p.add_(0.5 + 0.3 * torch.randn_like(p))
    twin = copy.deepcopy(enc)
    rng = np.random.RandomState(1)
    xs = torch.from_numpy(rng.randn(3, 37, enc.in_channel * enc.input_freq).astype(np.float32))
    y_ref = _general_stack(twin, xs, 1.7)
    w = torch.from_numpy(rng.randn(*y_ref.shape).astype(np.float32))
    (y_ref * w).sum().backward()
    ref = {k: p.grad.clone() for k, p in twin.named_parameters()}
    y = real_node(enc, xs, 1.7, "fp32")
    assert float((y - y_ref).abs().max()) <= 1e-5 * float(y_ref.abs().max())
    (y * w).sum().backward()
    gmax = max(float(v.abs().max()) for v in ref.values())
    for k, p in enc.named_parameters():
        d = (p.grad - ref[k]).abs() / max(float(ref[k].abs().max()), 1e-2 * gmax)
        assert int((d > 3e-3).sum()) <= 2 and float(d.max()) <= 5e-2, (k, float(d.max()), int((d > 3e-3).sum()))


def _bn2d_stack(enc, xs, out_scale):
    """torch restatement of the BatchNorm2d front-end in train() mode (reference conv.py:362-394); updates the modules' running
    statistics like the reference does."""
    F_ = torch.nn.functional
    B, T, Fd = xs.shape
    x = xs.view(B, T, enc.in_channel, Fd // enc.in_channel).transpose(1, 2)
    for blk in enc.layers:
        x = torch.relu(blk.norm1(F_.conv2d(x, blk.conv1.weight, blk.conv1.bias, padding=1)))
        x = torch.relu(blk.norm2(F_.conv2d(x, blk.conv2.weight, blk.conv2.bias, padding=1, stride=tuple(blk.stride))))
        if blk.pool is not None:
            x = F_.max_pool2d(x, blk.pooling, blk.pooling, ceil_mode=True)
    B, C, T, Fq = x.shape
    x = x.transpose(1, 2).reshape(B, T, C * Fq)
    if enc.bridge is not None:
        x = F_.linear(x, enc.bridge.weight, enc.bridge.bias)
    return x * out_scale


@pytest.mark.parametrize("ov", [
    dict(), dict(poolings="(1,1)_(2,2)", bottleneck_dim=24), dict(strides="(1,1)_(2,2)", poolings="(1,1)_(1,1)"),
    dict(input_dim=120, in_channel=3, channels="16_32"),
])
def test_batchnorm2d_frontend_node_matches_torch_autograd(ov, monkeypatch):
    """BatchNorm2d blocks in training: batch statistics over all B*T*F positions, the backward through the statistics
    (nsp_bn_bwd), and nn.BatchNorm2d's running-statistics update."""
    import copy
    import ops_doubles
    from neural_sp_b200 import autograd as ag
    from neural_sp_b200.encoders.conv import ConvEncoder
    real_node = ag.frontend_forward
    ops_doubles.install_training(monkeypatch)
    torch.manual_seed(0)
    enc = ConvEncoder(**_cfg(normalization='batch_norm', **ov)).train()
    enc.set_precision("fp32")
    with torch.no_grad():
        for blk in enc.layers:
            for n in (blk.norm1, blk.norm2):
                n.weight.add_(0.3 * torch.randn_like(n.weight))
                n.bias.add_(0.3 * torch.randn_like(n.bias))
    twin = copy.deepcopy(enc)
    rng = np.random.RandomState(1)
    xs = torch.from_numpy(rng.randn(3, 37, enc.in_channel * enc.input_freq).astype(np.float32))
    y_ref = _bn2d_stack(twin, xs, 1.7)
    w = torch.from_numpy(rng.randn(*y_ref.shape).astype(np.float32))
    (y_ref * w).sum().backward()
    ref = {k: p.grad.clone() for k, p in twin.named_parameters()}
    y = real_node(enc, xs, 1.7, "fp32")
    assert float((y - y_ref).abs().max()) <= 1e-5 * float(y_ref.abs().max())
    (y * w).sum().backward()
    gmax = max(float(v.abs().max()) for v in ref.values())
    for k, p in enc.named_parameters():
        # (a conv bias in front of a BatchNorm has NO gradient -- the mean is subtracted -- so its reference gradient is rounding
        # noise: errors are measured against max(|g|, 1e-2 of the largest gradient))
        d = (p.grad - ref[k]).abs() / max(float(ref[k].abs().max()), 1e-2 * gmax)
        assert int((d > 1e-3).sum()) <= 2 and float(d.max()) <= 5e-2, (k, float(d.max()), int((d > 1e-3).sum()))
    for (k, b), (_, bt) in zip(enc.named_buffers(), twin.named_buffers()):       # running_mean / running_var / num_batches_tracked
        assert torch.allclose(b.float(), bt.float(), rtol=1e-5, atol=1e-6), k
