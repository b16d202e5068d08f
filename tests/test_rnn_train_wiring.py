"""Host logic of the RNN encoder's training path on CPU: the autograd orchestration (neural_sp_b200/autograd.py
_LstmLayerFn / _LinearReluFn, flat gradient buckets, direction-fused views, subsamplers, sub-task outputs) is run with the
library's ops replaced by torch restatements (test doubles), and every parameter gradient is compared with the
UNMODIFIED reference's autograd (tests/golden/rnngrad_*.npz).  The CUDA kernels behind the real ops are checked by
tests/test_rnn_gpu.py; this test pins what Python does with their results."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load_golden

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from gen_golden_encoder import grad_loss_weights  # noqa: E402


def _linear(x, wp, bias=None, prec="fp32", act=None, glu=False, residual=None, alpha=1.0, out_dtype=torch.float32, out=None, **kw):
    y = x.float().reshape(-1, x.shape[-1]) @ wp[0].t()
    if bias is not None:
        y = y + bias
    if act == "relu":
        y = torch.relu(y)
    y = alpha * y
    if residual is not None:
        y = y + residual.reshape(y.shape)
    if out is not None:
        out.copy_(y.reshape(out.shape))
        return out
    return y.reshape(*x.shape[:-1], -1)


def _lstm_seq(gx, whh, lens, nd, save=False, prec=None):
    B, T, G = gx.shape
    H = G // (4 * nd)
    y, acts = torch.zeros(B, T, nd * H), torch.zeros(B, T, nd, 4 * H)
    cp, hp = torch.zeros(B, T, nd, H), torch.zeros(B, T, nd, H)
    for d in range(nd):
        h, c = torch.zeros(B, H), torch.zeros(B, H)
        for s in range(T):
            for b in range(B):
                n = int(lens[b])
                if s < n:
                    t = s if d == 0 else n - 1 - s
                    g = gx[b, t, d * 4 * H:(d + 1) * 4 * H] + whh[d] @ h[b]
                    i, f, o = torch.sigmoid(g[:H]), torch.sigmoid(g[H:2 * H]), torch.sigmoid(g[3 * H:])
                    gg = torch.tanh(g[2 * H:3 * H])
                    acts[b, t, d], cp[b, t, d], hp[b, t, d] = torch.cat([i, f, gg, o]), c[b], h[b]
                    c[b] = f * c[b] + i * gg
                    h[b] = o * torch.tanh(c[b])
                    y[b, t, d * H:(d + 1) * H] = h[b]
    return (y, acts, cp, hp) if save else y


def _lstm_seq_bwd(dy, acts, cprev, whh, lens, prec=None):
    """Same step structure as the kernel (csrc/lstm.cu): cell backward -> dG_t, then dh_rec = dG_t W_hh."""
    B, T, nd, H4 = acts.shape
    H = H4 // 4
    dG = torch.zeros(B, T, nd * H4)
    for d in range(nd):
        dc, dhr = torch.zeros(B, H), torch.zeros(B, H)
        for s in range(T - 1, -1, -1):
            X = torch.zeros(B, H4)
            for b in range(B):
                n = int(lens[b])
                if s < n:
                    t = s if d == 0 else n - 1 - s
                    a = acts[b, t, d]
                    i, f, gg, o = a[:H], a[H:2 * H], a[2 * H:3 * H], a[3 * H:]
                    c0 = cprev[b, t, d]
                    tc = torch.tanh(f * c0 + i * gg)
                    dh = dy[b, t, d * H:(d + 1) * H] + dhr[b]
                    dcc = dc[b] + dh * o * (1 - tc * tc)
                    dc[b] = dcc * f
                    X[b] = torch.cat([dcc * gg * i * (1 - i), dcc * c0 * f * (1 - f), dcc * i * (1 - gg * gg), dh * tc * o * (1 - o)])
                    dG[b, t, d * H4:(d + 1) * H4] = X[b]
            dhr = X @ whh[d]
    return dG


def _linear_wgrad(dy, x, prec, dw, alpha=1.0, accumulate=True):
    N, K = dw.shape
    v = alpha * dy.reshape(-1, N).float().t() @ x.reshape(-1, x.shape[-1])[:, :K].float()
    dw.copy_(dw + v if accumulate else v)
    return dw


def _colsum_acc(x, y, alpha=1.0):
    y.add_(alpha * x.reshape(-1, x.shape[-1]).float().sum(0))
    return y


def _pool_time(x, f, mode):
    assert mode == "max"
    return torch.nn.functional.max_pool1d(x.transpose(1, 2), f, f, ceil_mode=True).transpose(1, 2).contiguous()


def _maxpool_time_bwd(x, dy, f):
    with torch.enable_grad():
        xx = x.detach().clone().requires_grad_(True)
        _pool_time(xx, f, "max").backward(dy)
    return xx.grad


def _frontend_forward(enc, xs, out_scale, prec):
    """torch restatement of the Conv2dBlock stack without bridge (reference conv.py:347-396)."""
    B, T, Fd = xs.shape
    x = xs.view(B, T, enc.in_channel, Fd // enc.in_channel).transpose(1, 2)
    for blk in enc.layers:
        x = torch.relu(torch.nn.functional.conv2d(x, blk.conv1.weight, blk.conv1.bias, padding=1))
        x = torch.relu(torch.nn.functional.conv2d(x, blk.conv2.weight, blk.conv2.bias, padding=1))
        if blk.pool is not None:
            x = torch.nn.functional.max_pool2d(x, blk.pooling, blk.pooling, ceil_mode=True)
    assert enc.bridge is None
    B, C, T, F = x.shape
    return x.transpose(1, 2).reshape(B, T, C * F) * out_scale


@pytest.fixture
def torch_ops(monkeypatch):
    from neural_sp_b200 import ops, autograd as ag
    import neural_sp_b200.encoders.rnn as rnn_mod
    doubles = dict(prepare_weight=lambda w, prec: (w.float().contiguous(), None), linear=_linear, lstm_seq=_lstm_seq,
                   lstm_seq_bwd=_lstm_seq_bwd, linear_wgrad=_linear_wgrad, colsum_acc=_colsum_acc,
                   relu_mask=lambda dx, a: torch.where(a > 0, dx, torch.zeros_like(dx)), pool_time=_pool_time,
                   maxpool_time_bwd=_maxpool_time_bwd)
    for k, v in doubles.items():
        monkeypatch.setattr(ops, k, v)
    monkeypatch.setattr(ag, "frontend_forward", _frontend_forward)
    monkeypatch.setattr(rnn_mod, "lens_to_device", lambda xlens, dev: xlens.clone())


@pytest.mark.parametrize("name", ["rnn_blstm_sum.npz", "rnn_conv_lstm_proj.npz"])
def test_rnn_training_wiring_matches_reference_gradients(name, torch_ops):
    from neural_sp_b200.encoders.conv import ConvEncoder
    from neural_sp_b200.encoders.rnn import RNNEncoder
    g, gg = load_golden(name), load_golden("rnngrad_" + name[4:])
    c = json.loads(str(g["cfg"]))
    a = dict(c["args"])
    a["frontend_conv"] = ConvEncoder(**c["conv"]) if c["conv"] else None
    enc = RNNEncoder(**a)
    enc.load_state_dict({k[3:]: torch.from_numpy(np.asarray(g[k]).astype(np.float32)) for k in g.files if k.startswith("sd.")})
    enc.train()
    enc.set_precision("fp32")
    out = enc(torch.from_numpy(g["xs"]), torch.IntTensor(g["xlens"].tolist()), task="all")
    ys = out["ys"]["xs"]
    assert [int(v) for v in out["ys"]["xlens"]] == g["xlens_out"].tolist()
    assert float((ys.detach() - torch.from_numpy(g["ys"])).abs().max()) <= 1e-5
    loss = (ys * torch.from_numpy(grad_loss_weights(tuple(ys.shape), out["ys"]["xlens"].tolist()))).sum()
    if out["ys_sub1"]["xs"] is not None:
        s1 = out["ys_sub1"]["xs"]
        loss = loss + (s1 * torch.from_numpy(grad_loss_weights(tuple(s1.shape), out["ys_sub1"]["xlens"].tolist(), seed=99))).sum()
    loss.backward()
    assert abs(float(loss.detach()) - float(gg["loss"])) <= 1e-4 * max(1.0, abs(float(gg["loss"])))
    for k, p in enc.named_parameters():
        ref = torch.from_numpy(gg["g." + k])
        assert p.grad is not None, k
        err = float((p.grad - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
        assert err <= 1e-4, (k, err)
