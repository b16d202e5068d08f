"""RNNEncoder (LSTM / BLSTM, persistent LSTM kernel + tcgen05 input projections) vs the unmodified reference's outputs
(tests/golden/rnn_*.npz: BASELINE configs[0] BLSTM 2x256, conv_lstm with projections / sub-task / bridge, summed BLSTM
with concat subsampling).  fp32 mode 2e-4 of max|ref| (the cuDNN/CPU LSTM of the reference is fp32), bf16 mode 5e-2."""
import json

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
CASES = ["rnn_c1_blstm.npz", "rnn_conv_lstm_proj.npz", "rnn_blstm_sum.npz"]


def _build(g, precision):
    from neural_sp_b200.encoders.conv import ConvEncoder
    from neural_sp_b200.encoders.rnn import RNNEncoder
    c = json.loads(str(g["cfg"]))
    a = dict(c["args"])
    a["frontend_conv"] = ConvEncoder(**c["conv"]) if c["conv"] else None
    enc = RNNEncoder(**a)
    sd = {k[3:]: torch.from_numpy(np.asarray(g[k]).astype(np.float32)) for k in g.files if k.startswith("sd.")}
    assert set(enc.state_dict()) == set(sd)
    enc.load_state_dict(sd, strict=True)
    enc = enc.cuda().eval()
    enc.set_precision(precision)
    return enc


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", CASES)
def test_rnn_encoder_matches_reference(name, precision):
    g = load_golden(name)
    enc = _build(g, precision)
    out = enc(torch.from_numpy(g["xs"]).cuda(), torch.IntTensor(g["xlens"].tolist()), task="all")
    ys = out["ys"]["xs"].float().cpu().numpy()
    assert [int(v) for v in out["ys"]["xlens"]] == g["xlens_out"].tolist()
    assert ys.shape == g["ys"].shape
    tol = 2e-4 if precision == "fp32" else 5e-2
    err = np.abs(ys - g["ys"]).max() / np.abs(g["ys"]).max()
    assert err <= tol, (name, precision, err)
    if "ys_sub1" in g.files:
        s = out["ys_sub1"]["xs"].float().cpu().numpy()
        assert np.abs(s - g["ys_sub1"]).max() / np.abs(g["ys_sub1"]).max() <= tol
    # packed-sequence semantics: frames beyond each utterance's length are exactly zero before the bridge
    if "bridge.weight" not in [k[3:] for k in g.files]:
        for b, n in enumerate(g["xlens_out"].tolist()):
            assert np.all(ys[b, n:] == 0)


def test_lstm_kernel_vs_torch_lstm_large_hidden():
    """config-4 shaped layer (uni-LSTM 1024 units, B=32, T'=250 -> 60 here) against torch.nn.LSTM on the same GPU."""
    from neural_sp_b200 import ops
    torch.manual_seed(0)
    B, T, I, H = 32, 60, 320, 1024
    lstm = torch.nn.LSTM(I, H, 1, batch_first=True).cuda()
    x = torch.randn(B, T, I, device="cuda")
    lens = torch.tensor([T - (b % 7) * 3 for b in range(B)], dtype=torch.int32)
    lens, _ = lens.sort(descending=True)
    with torch.no_grad():
        packed = torch.nn.utils.rnn.pack_padded_sequence(x, lens.tolist(), batch_first=True)
        ref, _ = lstm(packed)
        ref = torch.nn.utils.rnn.pad_packed_sequence(ref, batch_first=True, total_length=T)[0]
        gx = torch.nn.functional.linear(x.double(), lstm.weight_ih_l0.double(), (lstm.bias_ih_l0 + lstm.bias_hh_l0).double()).float()
        y = ops.lstm_seq(gx, lstm.weight_hh_l0[None], lens.cuda(), 1)
    assert (y - ref).abs().max().item() <= 2e-4


# ---------------------------------------------------------------------------------------------
# training path: persistent BPTT kernel + GEMM gradients
# ---------------------------------------------------------------------------------------------
@pytest.fixture
def cudnn_fp32():
    """The comparison partner (cuDNN LSTM) must compute in fp32 too."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("shape", [(5, 23, 40, 64, 2), (32, 40, 80, 256, 2), (40, 12, 64, 1024, 1), (3, 1, 16, 8, 2)])
def test_lstm_bwd_kernel_vs_torch_autograd(shape, cudnn_fp32):
    """d loss / d gate pre-activations of nsp_lstm_seq_bwd, checked through everything torch exposes: dx, dW_ih, dW_hh and
    the bias gradients of nn.LSTM on packed (ragged) sequences, same GPU, fp32."""
    from neural_sp_b200 import ops
    B, T, I, H, nd = shape
    torch.manual_seed(1)
    lstm = torch.nn.LSTM(I, H, 1, batch_first=True, bidirectional=(nd == 2)).cuda()
    x = torch.randn(B, T, I, device="cuda", requires_grad=True)
    lens = torch.tensor(sorted([max(1, T - (b * 5) % T) for b in range(B)], reverse=True), dtype=torch.int32)
    packed = torch.nn.utils.rnn.pack_padded_sequence(x, lens.tolist(), batch_first=True)
    ref = torch.nn.utils.rnn.pad_packed_sequence(lstm(packed)[0], batch_first=True, total_length=T)[0]
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    sfx = ["_l0", "_l0_reverse"][:nd]
    with torch.no_grad():
        w_ih = torch.cat([getattr(lstm, "weight_ih" + s) for s in sfx]).double()
        bias = torch.cat([getattr(lstm, "bias_ih" + s) + getattr(lstm, "bias_hh" + s) for s in sfx]).double()
        whh = torch.stack([getattr(lstm, "weight_hh" + s) for s in sfx]).contiguous()
        gx = torch.nn.functional.linear(x.detach().double(), w_ih, bias).float()
        y, acts, cprev, hprev = ops.lstm_seq(gx, whh, lens.cuda(), nd, save=True)
        assert (y - ref).abs().max().item() <= 2e-4
        dg = ops.lstm_seq_bwd(w, acts, cprev, whh, lens.cuda()).double()
        for b, n in enumerate(lens.tolist()):
            assert torch.all(dg[b, n:] == 0)

        def close(got, want, what):
            err = (got - want.double()).abs().max().item() / max(want.abs().max().item(), 1e-30)
            assert err <= 1e-3, (what, err)
        close(dg @ w_ih, x.grad, "dx")
        for d, s in enumerate(sfx):
            g = dg[:, :, d * 4 * H:(d + 1) * 4 * H].reshape(-1, 4 * H)
            close(g.t() @ x.detach().double().reshape(-1, I), getattr(lstm, "weight_ih" + s).grad, "dW_ih" + s)
            close(g.t() @ hprev[:, :, d].double().reshape(-1, H), getattr(lstm, "weight_hh" + s).grad, "dW_hh" + s)
            close(g.sum(0), getattr(lstm, "bias_ih" + s).grad, "db" + s)


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-3), ("bf16", 2e-1)])
@pytest.mark.parametrize("name", ["rnn_conv_lstm_proj.npz", "rnn_blstm_sum.npz"])
def test_rnn_encoder_param_grads_match_reference(name, precision, tol):
    """train() + grad mode: every parameter gradient against the unmodified reference's autograd (rnngrad_*.npz)."""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from gen_golden_encoder import grad_loss_weights
    g = load_golden(name)
    gg = load_golden("rnngrad_" + name[4:])
    enc = _build(g, precision).train()
    out = enc(torch.from_numpy(g["xs"]).cuda(), torch.IntTensor(g["xlens"].tolist()), task="all")
    ys = out["ys"]["xs"]
    assert ys.requires_grad
    loss = (ys * torch.from_numpy(grad_loss_weights(tuple(ys.shape), out["ys"]["xlens"].tolist())).cuda()).sum()
    if out["ys_sub1"]["xs"] is not None:
        s1 = out["ys_sub1"]["xs"]
        loss = loss + (s1 * torch.from_numpy(grad_loss_weights(tuple(s1.shape), out["ys_sub1"]["xlens"].tolist(), seed=99)).cuda()).sum()
    loss.backward()
    gmax = max(float(np.abs(gg[k]).max()) for k in gg.files if k.startswith("g."))
    errs = {}
    for k, p in enc.named_parameters():
        ref = gg["g." + k]
        assert p.grad is not None, k
        errs[k] = float(np.abs(p.grad.float().cpu().numpy() - ref).max() / max(float(np.abs(ref).max()), 1e-3 * gmax))
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    assert worst[0][1] <= tol, (name, precision, worst)
