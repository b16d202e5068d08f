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
