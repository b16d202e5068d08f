"""RNN-T loss (nsp_rnnt_loss_fwd_bwd) vs the torchaudio goldens, the oracle DP, and the reference's stand-in op
on the GPU box at config-4 scale.  Tolerance: nll 1e-4 relative, gradient 2e-4 absolute (entries <= 1/B)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["rnnt_small.npz", "rnnt_mid.npz"])
def test_rnnt_matches_golden(name):
    from neural_sp_b200 import ops
    g = load_golden(name)
    dev = "cuda"
    lp = torch.from_numpy(g["logits"]).to(dev).log_softmax(-1)
    loss, nll, grad = ops.rnnt_loss_fwd_bwd(lp, torch.from_numpy(g["ys"]).to(dev), torch.from_numpy(g["flens"]).to(dev),
                                            torch.from_numpy(g["ylens"]).to(dev), 0)
    np.testing.assert_allclose(nll.cpu().numpy(), g["nll"], rtol=1e-4)
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    np.testing.assert_allclose(grad.cpu().numpy(), g["grad_log_probs"], atol=2e-4, rtol=0)


def test_rnnt_vs_oracle_ragged():
    from neural_sp_b200 import ops
    from oracle import rnnt_oracle
    rng = np.random.default_rng(2)
    torch.manual_seed(2)
    B, T, U, V = 4, 33, 9, 17
    lp = (torch.randn(B, T, U + 1, V, device="cuda") * 2).log_softmax(-1)
    flens = np.array([33, 20, 1, 12], np.int32)
    ylens = np.array([9, 3, 0, 1], np.int32)
    ys = rng.integers(1, V, size=(B, U)).astype(np.int32)
    loss, nll, grad = ops.rnnt_loss_fwd_bwd(lp, torch.from_numpy(ys).cuda(), torch.from_numpy(flens).cuda(),
                                            torch.from_numpy(ylens).cuda(), 0)
    o_nll, o_loss, o_grad = rnnt_oracle.rnnt_nll_and_grad(lp.cpu().numpy(), ys, flens, ylens)
    np.testing.assert_allclose(nll.cpu().numpy(), o_nll, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(grad.cpu().numpy(), o_grad, atol=2e-4, rtol=0)
    assert abs(loss.item() - o_loss) <= 1e-4 * abs(o_loss)


def test_rnnt_config4_scale_vs_torchaudio():
    """C4 offline shape (B=8 slice of 32, T'=250, U=56, V=1000): compare with torchaudio's rnnt_loss on the GPU."""
    import torchaudio
    from neural_sp_b200 import ops
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    B, T, U, V = 8, 250, 56, 1000
    logits = torch.randn(B, T, U + 1, V, device="cuda")
    lp = logits.log_softmax(-1).requires_grad_(True)
    flens = torch.tensor([250 - 7 * b for b in range(B)], dtype=torch.int32, device="cuda")
    ylens = torch.tensor([56 - 3 * b for b in range(B)], dtype=torch.int32, device="cuda")
    ys = torch.from_numpy(rng.integers(1, V, size=(B, U)).astype(np.int32)).cuda()
    ref = torchaudio.functional.rnnt_loss(lp, ys, flens, ylens, blank=0, reduction="none", fused_log_softmax=False)
    ref.mean().backward()
    loss, nll, grad = ops.rnnt_loss_fwd_bwd(lp.detach(), ys, flens, ylens, 0)
    assert torch.allclose(nll, ref.detach(), rtol=1e-4, atol=1e-2)
    assert (grad - lp.grad).abs().max().item() <= 2e-4
    # property: every cell's outgoing probability mass is conserved -> sum over the blank column at t = T-1, u = U is -1/B
    assert abs(grad[0, int(flens[0]) - 1, int(ylens[0]), 0].item() + 1.0 / B) < 5e-4    # fp32 lattice noise at T=250


def test_rnn_transducer_module_matches_torch_chain():
    """RNNTransducer.forward_transducer (joint GEMMs + tanh + log-softmax + lattice kernel) vs the same chain in plain
    torch fp32 with torchaudio's loss as the stand-in for warp_rnnt (rnn_transducer.py:236-258)."""
    import torchaudio
    from neural_sp_b200.decoders.rnn_transducer import RNNTransducer
    torch.manual_seed(0)
    sym = {'eos': 2, 'unk': 1, 'pad': 3, 'blank': 0}
    dec = RNNTransducer(sym, enc_n_units=48, n_units=32, n_projs=0, n_layers=2, bottleneck_dim=40, emb_dim=16, vocab=50,
                        dropout=0.0, dropout_emb=0.0, ctc_weight=0.0, ctc_lsm_prob=0.0, ctc_fc_list="", external_lm=None,
                        global_weight=1.0, mtl_per_batch=False, param_init=0.1).cuda().eval()
    dec.set_precision("fp32")
    B, T = 3, 30
    eouts = torch.randn(B, T, 48, device="cuda")
    elens = torch.IntTensor([30, 25, 18])
    ys = [[5, 6, 7, 8, 9], [10, 11, 12], [4]]
    loss = dec.forward_transducer(eouts, elens, ys)
    # reference chain
    U = 5
    ys_in = torch.full((B, U + 1), 3, dtype=torch.long)
    ys_out = torch.zeros(B, U, dtype=torch.int32)
    for b, y in enumerate(ys):
        ys_in[b, 0] = 2
        ys_in[b, 1:len(y) + 1] = torch.tensor(y)
        ys_out[b, :len(y)] = torch.tensor(y, dtype=torch.int32)
    with torch.no_grad():
        d = dec.embed(ys_in.cuda())
        for l in range(2):
            d, _ = dec.rnn[l](d)
        z = torch.tanh(dec.w_enc(eouts)[:, :, None] + dec.w_dec(d)[:, None])
        lp = torch.log_softmax(dec.output(z), -1)
        ref = torchaudio.functional.rnnt_loss(lp, ys_out.cuda(), elens.cuda(), torch.tensor([5, 3, 1], dtype=torch.int32).cuda(),
                                              blank=0, reduction="mean", fused_log_softmax=False)
    assert loss.shape == (1,)
    assert abs(loss.item() - ref.item()) <= 1e-3 * abs(ref.item()), (loss.item(), ref.item())
