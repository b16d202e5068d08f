"""Parity of the CUDA CTC path (through the C ABI) against the reference's golden vectors, the
oracle on seeded inputs, and -- at BASELINE.json's full sizes -- size-independent properties plus
the reference's own op (torch.nn.functional.ctc_loss) run on the same GPU.

Tolerances: loss 1e-4 relative (north_star asks 1e-3), gradient 2e-4 absolute on values <= 1/B
(the reference itself is fp32), trigger points bit-exact.
"""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden, split_labels

pytestmark = pytest.mark.gpu

CTC_CASES = sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLDEN, "ctc_*.npz")) if "head" not in f)


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", CTC_CASES)
def test_ctc_loss_matches_reference_golden(name):
    from neural_sp_b200 import ops
    g = load_golden(name)
    dev = _dev()
    ys = split_labels(g["ys_cat"], g["ylens"])
    logits = torch.from_numpy(g["logits"]).to(dev)
    labels, ylens, _ = ops.pack_labels(ys, dev)
    elens = torch.from_numpy(g["elens"]).to(dev)
    loss, nll, grad = ops.ctc_loss_fwd_bwd(logits, labels, elens, ylens, 0, float(g["lsm"]))
    torch.cuda.synchronize()
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    np.testing.assert_allclose(nll.cpu().numpy(), g["nll"], rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(grad.cpu().numpy(), g["grad"], rtol=0, atol=2e-4)


@pytest.mark.parametrize("name", [c for c in CTC_CASES if "trigger_points" in np.load(os.path.join(GOLDEN, c)).files])
def test_forced_align_bit_exact_vs_reference_golden(name):
    from neural_sp_b200 import ops
    g = load_golden(name)
    dev = _dev()
    ys = split_labels(g["ys_cat"], g["ylens"])
    labels, ylens, _ = ops.pack_labels(ys, dev)
    trig = ops.ctc_forced_align(torch.from_numpy(g["logits"]).to(dev), labels,
                                torch.from_numpy(g["elens"]).to(dev), ylens, 0)
    assert np.array_equal(trig.cpu().numpy(), g["trigger_points"])


def test_ctc_transposed_view_and_autograd():
    """loss_fn receives the [T,B,V] transpose view (ctc.py:125); backward scales the fused gradient."""
    from neural_sp_b200.decoders.ctc import CTC
    g = load_golden("ctc_las_test_shape.npz")
    dev = _dev()
    logits = torch.from_numpy(g["logits"]).to(dev).requires_grad_(True)
    ctc = CTC(eos=2, blank=0, enc_n_units=8, vocab=logits.size(2), lsm_prob=0.0).to(dev)
    loss = ctc.loss_fn(logits.transpose(1, 0), torch.from_numpy(g["ys_cat"]), torch.from_numpy(g["elens"]),
                       torch.from_numpy(g["ylens"]))
    (2.0 * loss).backward()
    assert abs(loss.item() - float(g["loss_ctc"])) <= 1e-4 * abs(float(g["loss_ctc"]))
    # golden grad includes lsm mixing; compare against the oracle instead for the pure-CTC gradient
    from oracle import ctc_oracle
    ys = split_labels(g["ys_cat"], g["ylens"])
    _, _, og = ctc_oracle.ctc_nll_and_grad(g["logits"], ys, g["elens"])
    np.testing.assert_allclose(logits.grad.cpu().numpy(), 2.0 * og, rtol=0, atol=2e-4)


def test_ctc_head_forward_matches_reference_golden():
    """CTC.forward through the fc head (ctc.py:105-137): loss, d loss/d eouts, parameter grads, triggers."""
    from neural_sp_b200.decoders.ctc import CTC
    g = load_golden("ctc_head_forward.npz")
    dev = _dev()
    D = g["eouts"].shape[2]
    V = g["sd.output.fc1.weight"].shape[0]
    ctc = CTC(eos=2, blank=0, enc_n_units=D, vocab=V, lsm_prob=0.1, fc_list="16").to(dev)
    ctc.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")})
    ctc.set_precision("fp32")          # parity mode: 3xTF32 GEMMs in the head (forward and backward)
    ctc.train()
    eouts = torch.from_numpy(g["eouts"]).to(dev).requires_grad_(True)
    ys = split_labels(g["ys_cat"], g["ylens"])
    loss, trig = ctc(eouts, torch.from_numpy(g["elens"]), ys, forced_align=True)
    loss.backward()
    assert loss.dim() == 0
    assert abs(loss.item() - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))
    assert np.array_equal(trig.cpu().numpy(), g["trigger_points"])
    ref = g["grad_eouts"]
    err = np.abs(eouts.grad.cpu().numpy() - ref).max()
    assert err <= 1e-3 * np.abs(ref).max() + 1e-6, err
    for n, p_ in ctc.named_parameters():
        r = g["gsd." + n]
        assert np.abs(p_.grad.cpu().numpy() - r).max() <= 1e-3 * np.abs(r).max() + 1e-6, n


def _synthetic(B, T, V, seed, ragged=True):
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    logits = torch.randn(B, T, V, device="cuda") * 2.0
    elens = np.full(B, T, np.int32)
    if ragged:
        elens = rng.integers(T // 2, T + 1, size=B).astype(np.int32)
        elens[0] = T
    ylens = np.minimum((0.45 * elens).astype(np.int32), elens)       # BASELINE.md label rule
    ys = [rng.integers(4, V, size=int(n)).tolist() for n in ylens]
    return logits, elens, ys


@pytest.mark.parametrize("B,T,V", [(32, 125, 1000), (32, 125, 10000), (32, 250, 10000), (3, 17, 50257), (2, 50, 13001)])
def test_ctc_full_size_properties_and_reference_op(B, T, V):
    from neural_sp_b200 import ops
    logits, elens, ys = _synthetic(B, T, V, seed=B + T + V)
    dev = logits.device
    labels, ylens_d, _ = ops.pack_labels(ys, dev)
    elens_d = torch.from_numpy(elens).to(dev)
    lsm = 0.1
    loss, nll, grad = ops.ctc_loss_fwd_bwd(logits, labels, elens_d, ylens_d, 0, lsm)
    # (1) rows of the gradient sum to zero (softmax and state occupancies both sum to one); pads are zero
    rs = grad.sum(-1)
    assert rs.abs().max().item() < 2e-4, rs.abs().max().item()
    for b in range(B):
        assert torch.count_nonzero(grad[b, int(elens[b]):]).item() == 0
    # (2) the reference's own op on the same GPU: log_softmax -> ctc_loss(sum, zero_infinity) / B (+ KL)
    x = logits.clone().requires_grad_(True)
    ys_cat = torch.tensor([v for y in ys for v in y], dtype=torch.int32)
    ref_nll = torch.nn.functional.ctc_loss(x.transpose(0, 1).log_softmax(2), ys_cat, torch.from_numpy(elens),
                                           torch.tensor([len(y) for y in ys], dtype=torch.int32),
                                           reduction="none", zero_infinity=True)
    lp = x.log_softmax(-1)
    mask = (torch.arange(T, device=dev)[None, :] < elens_d[:, None]).unsqueeze(-1)
    kl = (lp.exp() * (lp - np.log(1.0 / (V - 1))) * mask).sum() / float(elens.sum())
    ref_loss = ref_nll.sum() / B * (1 - lsm) + kl * lsm
    ref_loss.backward()
    assert torch.allclose(nll, ref_nll.detach(), rtol=1e-4, atol=1e-2)
    assert abs(loss.item() - ref_loss.item()) <= 1e-4 * abs(ref_loss.item())
    # compare on valid frames only: torch's CUDA ctc_loss backward leaves non-zero values in rows
    # t >= input_length (the CPU path the goldens come from, and this library, give exact zeros there)
    valid = mask.expand_as(grad)
    assert ((grad - x.grad) * valid).abs().max().item() <= 2e-4


def test_ctc_many_labels_spt_paths():
    """2L+1 > 512 exercises the multi-state-per-thread lattice variants."""
    from neural_sp_b200 import ops
    from oracle import ctc_oracle
    rng = np.random.default_rng(5)
    B, T, V = 2, 700, 40
    torch.manual_seed(3)
    logits = torch.randn(B, T, V, device="cuda")
    ys = [rng.integers(1, V, size=300).tolist(), rng.integers(1, V, size=17).tolist()]
    elens = np.array([700, 650], np.int32)
    labels, ylens_d, _ = ops.pack_labels(ys, logits.device)
    loss, nll, grad = ops.ctc_loss_fwd_bwd(logits, labels, torch.from_numpy(elens).cuda(), ylens_d, 0, 0.0)
    o_nll, o_loss, o_grad = ctc_oracle.ctc_nll_and_grad(logits.cpu().numpy(), ys, elens)
    np.testing.assert_allclose(nll.cpu().numpy(), o_nll, rtol=2e-4)
    # T=700, L=300: an fp32 log-domain lattice carries ~sqrt(T)*ulp(1e3) ~ 1e-3 absolute noise (so does ATen's)
    np.testing.assert_allclose(grad.cpu().numpy(), o_grad, atol=2e-3, rtol=0)
    trig = ops.ctc_forced_align(logits, labels, torch.from_numpy(elens).cuda(), ylens_d, 0)
    assert np.array_equal(trig.cpu().numpy(), ctc_oracle.forced_align(logits.cpu().numpy(), elens, ys))


def test_ops_reject_cpu_tensors():
    from neural_sp_b200 import ops, _lib
    with pytest.raises(_lib.NspError):
        ops.ctc_loss_fwd_bwd(torch.zeros(1, 2, 4), torch.zeros(1, 1, dtype=torch.int32),
                             torch.ones(1, dtype=torch.int32), torch.ones(1, dtype=torch.int32))


def test_greedy_trigger_points_probs_scores():
    """CTC.greedy / trigger_points / probs / scores (ctc.py:152-243) vs the reference's formulas in plain torch."""
    from itertools import groupby
    from neural_sp_b200.decoders.ctc import CTC
    torch.manual_seed(0)
    dev = _dev()
    ctc = CTC(eos=2, blank=0, enc_n_units=24, vocab=12, lsm_prob=0.0).to(dev).eval()
    ctc.set_precision("fp32")
    eouts = torch.randn(3, 40, 24, device=dev) * 3
    elens = torch.IntTensor([40, 33, 9])
    with torch.no_grad():
        logits = torch.nn.functional.linear(eouts, ctc.output.weight, ctc.output.bias)
    best = logits.log_softmax(-1).argmax(-1).cpu()
    hyps = ctc.greedy(eouts, elens.numpy())
    trig = ctc.trigger_points(eouts, elens).cpu()
    for b in range(3):
        idx = [int(best[b, t]) for t in range(int(elens[b]))]
        ref = [x for x in (g[0] for g in groupby(idx)) if x != 0]
        assert hyps[b][0] == ref
        tp = [t for t in range(int(elens[b])) if idx[t] != 0 and (t == 0 or idx[t] != idx[t - 1])]
        assert trig[b, :len(tp)].tolist() == tp and trig[b, len(tp):].abs().sum().item() == 0
    p, s = ctc.probs(eouts, temperature=2.0), ctc.scores(eouts)
    assert torch.allclose(p, torch.softmax(logits / 2.0, -1), atol=1e-5)
    assert torch.allclose(s, torch.log_softmax(logits, -1), atol=2e-4)


# ---- the streaming kernel (csrc/ctc_stream.cuh): both tile modes, every lattice width, ragged / empty / infeasible
# ---- utterances, more utterances than SMs (several lattices per CTA), against the pinned oracle (fp64 CPU restatement)
STREAM_CASES = [
    # B,   T,   V,     Lmax, lsm, note
    (6,   33,  2000,  9,   0.1),     # ROW mode, 1 float4 slot per thread, K = 1
    (4,   64,  12288, 20,  0.1),     # ROW mode at its widest (6 slots), K = 2
    (3,   50,  1280,  40,  0.0),     # WARP mode at its widest (10 slots per lane), K = 4
    (2,   200, 32,    30,  0.0),     # BASELINE configs[0] shape: WARP mode, tiny rows, many rows per tile
    (200, 40,  320,   12,  0.1),     # more utterances than SMs: every CTA sweeps several lattices
    (2,   300, 64,    120, 0.1),     # K = 8
    (2,   460, 48,    200, 0.0),     # K = 16
    (5,   70,  5000,  30,  0.1),     # ROW mode, V4 = 1250 -> 3 slots, ragged with padded tiles
]


@pytest.mark.parametrize("B,T,V,Lmax,lsm", STREAM_CASES)
def test_ctc_stream_kernel_vs_oracle(B, T, V, Lmax, lsm):
    from neural_sp_b200 import ops
    from oracle import ctc_oracle
    rng = np.random.default_rng(B * 7 + T * 3 + V)
    torch.manual_seed(B + T + V)
    logits = (torch.randn(B, T, V, device="cuda") * 2.0).contiguous()
    elens = rng.integers(max(1, T // 3), T + 1, size=B).astype(np.int32)
    elens[0] = T
    ylens = np.minimum(rng.integers(1, Lmax + 1, size=B), np.maximum(elens // 2, 1)).astype(np.int32)
    ylens[0] = min(Lmax, T // 2)
    ys = [rng.integers(1, V, size=int(n)).tolist() for n in ylens]
    if B >= 3:
        ys[1] = [ys[1][0]] * len(ys[1])                     # all repeats: needs 2L-1 frames, same-label chains
        if 2 * len(ys[1]) - 1 > elens[1]:
            pass                                            # infeasible on purpose (zero_infinity path)
        elens[2] = 0                                        # empty utterance
    if B >= 5:
        ys[4] = []                                          # empty label sequence
    labels, ylens_d, _ = ops.pack_labels(ys, logits.device)
    loss, nll, grad = ops.ctc_loss_fwd_bwd(logits, labels, torch.from_numpy(elens).cuda(), ylens_d, 0, lsm)
    torch.cuda.synchronize()
    o_loss, o_grad, o_nll = ctc_oracle.ctc_forward(logits.cpu().numpy(), ys, elens, lsm)
    np.testing.assert_allclose(nll.cpu().numpy(), o_nll, rtol=2e-4, atol=2e-3)
    assert abs(loss.item() - o_loss) <= 2e-4 * max(1.0, abs(o_loss)), (loss.item(), o_loss)
    tol = 2e-4 if T < 200 else (4e-4 if T <= 250 else 1e-3)  # fp32 log-domain lattices carry ~sqrt(T) ulp(|log p|) of noise (so does ATen's)
    np.testing.assert_allclose(grad.cpu().numpy(), o_grad, atol=tol, rtol=0)
    for b in range(B):
        assert torch.count_nonzero(grad[b, int(elens[b]):]).item() == 0


def test_ctc_strided_logits_take_the_row_tiles():
    """[T,B,V]-major logits viewed as [B,T,V] (the reference's loss_fn boundary, ctc.py:139-150): rows are not contiguous
    tile to tile; V = 2000 runs the one-row-per-tile streaming mode with strided row addresses, V = 400 the register kernel."""
    from neural_sp_b200 import ops
    from oracle import ctc_oracle
    rng = np.random.default_rng(3)
    for V in (2000, 400):
        B, T = 4, 37
        torch.manual_seed(V)
        tbv = torch.randn(T, B, V, device="cuda") * 1.5
        logits = tbv.transpose(0, 1)
        elens = np.array([37, 30, 21, 9], np.int32)
        ys = [rng.integers(1, V, size=n).tolist() for n in (12, 9, 4, 3)]
        labels, ylens_d, _ = ops.pack_labels(ys, logits.device)
        loss, nll, grad = ops.ctc_loss_fwd_bwd(logits, labels, torch.from_numpy(elens).cuda(), ylens_d, 0, 0.1)
        o_loss, o_grad, o_nll = ctc_oracle.ctc_forward(logits.cpu().numpy(), ys, elens, 0.1)
        assert abs(loss.item() - o_loss) <= 2e-4 * abs(o_loss)
        np.testing.assert_allclose(grad.cpu().numpy(), o_grad, atol=2e-4, rtol=0)
