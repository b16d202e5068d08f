"""Tensor-core LSTM recurrence (lstm_tc.cu, bf16 operands / fp32 accumulation and state) against the fp32 CUDA-core
recurrence (lstm.cu, itself pinned to torch.nn.LSTM and the reference's goldens in test_rnn_gpu.py) on the same inputs:
forward outputs / saved activations / final state, and the BPTT kernel on IDENTICAL saved activations.
Bounds: |y| <= 1; the only difference is bf16 rounding of h_{t-1}, W_hh (forward) and dG_t, W_hh (backward) in the
K-long dot products -> 2e-2 absolute on y (measured ~3e-3), 3e-2 relative L2 on dG."""
import pytest
import torch

pytestmark = pytest.mark.gpu

#        B   T    H   nd  ragged
CASES = [(32, 40, 256, 2, True),      # configs[0] BLSTM layer
         (32, 30, 1024, 1, True),     # configs[3] uni-LSTM layer (128 CTAs forward, 64 backward)
         (5, 17, 64, 1, True),        # fewer rows than one swizzle atom, single k chunk
         (9, 23, 320, 2, False),      # H / 64 odd: one chunk per ring stage
         (70, 12, 128, 2, True),      # B > 64: M = 128 tiles
         (64, 9, 192, 1, False)]


def _inputs(B, T, H, nd, ragged, seed=0):
    g = torch.Generator().manual_seed(seed)
    gx = torch.randn(B, T, nd * 4 * H, generator=g).cuda()
    whh = ((torch.rand(nd, 4 * H, H, generator=g) * 2 - 1) / H ** 0.5).cuda()
    lens = torch.tensor([max(1, T - (b * 5) % T) if ragged else T for b in range(B)], dtype=torch.int32).cuda()
    return gx, whh, lens


@pytest.mark.parametrize("B,T,H,nd,ragged", CASES)
def test_forward_matches_fp32_recurrence(B, T, H, nd, ragged):
    from neural_sp_b200 import ops
    from neural_sp_b200._lib import lib
    assert lib.nsp_lstm_tc_supported(B, H, nd) == 1
    gx, whh, lens = _inputs(B, T, H, nd, ragged)
    y0, a0, c0, h0 = ops.lstm_seq(gx, whh, lens, nd, save=True)
    y1, a1, c1, h1 = ops.lstm_seq(gx, whh, lens, nd, save=True, prec="bf16")
    torch.cuda.synchronize()
    assert torch.isfinite(y1).all()
    assert (y1 - y0).abs().max().item() <= 2e-2, (y1 - y0).abs().max().item()
    assert (y1 - y0).abs().mean().item() <= 2e-3
    for got, want in ((a1, a0), (c1, c0), (h1, h0)):
        assert (got - want).abs().max().item() <= 5e-2
    for b, n in enumerate(lens.tolist()):                       # pad_packed_sequence semantics
        assert torch.all(y1[b, n:] == 0) and torch.all(a1[b, n:] == 0)


@pytest.mark.parametrize("upc", ["8", "16"])
@pytest.mark.parametrize("B,T,H,nd,ragged", CASES)
def test_backward_matches_fp32_recurrence(B, T, H, nd, ragged, upc, monkeypatch):
    """Both decompositions of the backward kernel: 8 units per CTA (default when H / 8 CTAs are co-resident) and 16."""
    from neural_sp_b200 import ops
    monkeypatch.setenv("NSP_LSTM_TC_BWD_UPC", upc)
    gx, whh, lens = _inputs(B, T, H, nd, ragged, seed=1)
    y, acts, cprev, hprev = ops.lstm_seq(gx, whh, lens, nd, save=True)
    dy = torch.randn_like(y)
    d0 = ops.lstm_seq_bwd(dy, acts, cprev, whh, lens)
    d1 = ops.lstm_seq_bwd(dy, acts, cprev, whh, lens, prec="bf16")
    torch.cuda.synchronize()
    assert torch.isfinite(d1).all()
    rel = ((d1 - d0).norm() / d0.norm()).item()
    assert rel <= 3e-2, rel
    for b, n in enumerate(lens.tolist()):
        assert torch.all(d1[b, n:] == 0)


@pytest.mark.parametrize("B,T,H", [(8, 21, 128), (32, 16, 512)])
def test_state_arguments(B, T, H):
    """h0 / c0 in, hN / cN out (streaming, rnn.py:343-346); dhN / dcN in, dh0 / dc0 out (LC-BLSTM training, rnn.py:454-498)."""
    from neural_sp_b200 import ops
    gx, whh, lens = _inputs(B, T, H, 1, True, seed=2)
    g = torch.Generator().manual_seed(3)
    h0 = (torch.randn(1, B, H, generator=g) * 0.5).cuda()
    c0 = (torch.randn(1, B, H, generator=g) * 0.5).cuda()
    ya, aa, ca, ha, (hNa, cNa) = ops.lstm_seq(gx, whh, lens, 1, save=True, state=(h0, c0), want_state=True)
    yb, ab, cb, hb, (hNb, cNb) = ops.lstm_seq(gx, whh, lens, 1, save=True, state=(h0, c0), want_state=True, prec="bf16")
    assert (yb - ya).abs().max().item() <= 2e-2
    assert (hNb - hNa).abs().max().item() <= 2e-2 and (cNb - cNa).abs().max().item() <= 5e-2
    # chunked == whole (the carried state is fp32; each chunk re-rounds only its operands)
    k = T // 2
    lk = lens.clamp(max=k)
    lr = (lens - k).clamp(min=0)
    y1, st = ops.lstm_seq(gx[:, :k].contiguous(), whh, lk, 1, state=(h0, c0), prec="bf16")
    assert (y1 - yb[:, :k]).abs().max().item() <= 1e-5
    if int(lr.min()) > 0:
        y2, _ = ops.lstm_seq(gx[:, k:].contiguous(), whh, lr, 1, state=st, prec="bf16")
        assert (y2 - yb[:, k:]).abs().max().item() <= 1e-5
    dy = torch.randn_like(ya)
    dhN = torch.randn(1, B, H, generator=g).cuda()
    dcN = torch.randn(1, B, H, generator=g).cuda()
    da, (dh0a, dc0a) = ops.lstm_seq_bwd(dy, aa, ca, whh, lens, dstate=(dhN, dcN), want_dstate=True)
    db, (dh0b, dc0b) = ops.lstm_seq_bwd(dy, aa, ca, whh, lens, dstate=(dhN, dcN), want_dstate=True, prec="bf16")
    assert ((db - da).norm() / da.norm()).item() <= 3e-2
    assert ((dh0b - dh0a).norm() / dh0a.norm()).item() <= 3e-2
    assert ((dc0b - dc0a).norm() / dc0a.norm()).item() <= 3e-2


def test_unsupported_shapes_fall_back():
    from neural_sp_b200 import ops
    from neural_sp_b200._lib import lib
    assert lib.nsp_lstm_tc_supported(32, 40, 1) == 0          # H % 64 != 0
    assert lib.nsp_lstm_tc_supported(129, 256, 1) == 0
    gx, whh, lens = _inputs(4, 6, 40, 1, False)
    y0 = ops.lstm_seq(gx, whh, lens, 1)
    y1 = ops.lstm_seq(gx, whh, lens, 1, prec="bf16")            # fp32 kernel
    assert torch.equal(y0, y1)
