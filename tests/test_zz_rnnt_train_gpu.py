"""RNN-T training path on the GPU: the new backward kernels (log-softmax backward with a device-side upstream scale, tanh
joint backward reductions) against torch autograd, and a whole `RNNTransducer` training step (prediction network on the
persistent LSTM kernels, joint + loss node) against the same chain in plain torch fp32 with torchaudio's loss as the stand-in
for warp_rnnt (rnn_transducer.py:236-258): loss, d loss / d encoder output and every parameter gradient.
(File name sorts last on purpose: written after the round's GPU budget was spent; the host wiring is pinned to the
unmodified reference on CPU by tests/test_rnnt_train_wiring_cpu.py.)"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,V", [(37, 50), (5, 1000), (3, 2500), (64, 8)])
def test_log_softmax_bwd(rows, V):
    from neural_sp_b200 import ops
    torch.manual_seed(V)
    z = torch.randn(rows, V, device="cuda", requires_grad=True)
    lp = torch.log_softmax(z, -1)
    dlp = torch.randn(rows, V, device="cuda") * (torch.rand(rows, V, device="cuda") < 0.1)
    g = torch.tensor(0.7, device="cuda")
    (lp * dlp).sum().mul(g).backward()
    out = ops.log_softmax_bwd_(lp.detach().contiguous(), dlp.clone(), g)
    assert torch.allclose(out, z.grad, atol=2e-6, rtol=1e-4)
    out1 = ops.log_softmax_bwd_(lp.detach().contiguous(), dlp.clone())
    assert torch.allclose(out1 * 0.7, z.grad, atol=2e-6, rtol=1e-4)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("B,T,U1,J", [(2, 7, 4, 32), (3, 30, 6, 40), (1, 1, 1, 640)])
def test_rnnt_joint_tanh_bwd(dtype, tol, B, T, U1, J):
    from neural_sp_b200 import ops
    torch.manual_seed(T)
    e = torch.randn(B, T, J, device="cuda", requires_grad=True)
    d = torch.randn(B, U1, J, device="cuda", requires_grad=True)
    h = torch.tanh(e[:, :, None] + d[:, None])
    dh = torch.randn_like(h)
    h.backward(dh)
    hk = ops.rnnt_joint_tanh(e.detach(), d.detach(), out_dtype=dtype)
    assert torch.allclose(hk.float(), h.detach(), atol=1e-2 if dtype == torch.bfloat16 else 1e-6)
    de, dd = ops.rnnt_joint_tanh_bwd(hk, dh.to(dtype))
    for got, ref in ((de, e.grad), (dd, d.grad)):
        assert float((got - ref).abs().max()) <= tol * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("n_projs", [0, 24])
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-3), ("bf16", 1e-1)])
def test_rnn_transducer_training_step(n_projs, precision, tol):
    import torchaudio
    from neural_sp_b200.decoders.rnn_transducer import RNNTransducer
    torch.manual_seed(0)
    sym = {'eos': 2, 'unk': 1, 'pad': 3, 'blank': 0}
    dec = RNNTransducer(sym, enc_n_units=48, n_units=32, n_projs=n_projs, n_layers=2, bottleneck_dim=40, emb_dim=16, vocab=56,
                        dropout=0.0, dropout_emb=0.0, ctc_weight=0.0, ctc_lsm_prob=0.0, ctc_fc_list="", external_lm=None,
                        global_weight=1.0, mtl_per_batch=False, param_init=0.1).cuda().train()
    dec.set_precision(precision)
    B, T = 3, 30
    e0 = torch.randn(B, T, 48, device="cuda")
    elens = torch.IntTensor([30, 25, 18])
    ys = [[5, 6, 7, 8, 9], [10, 11, 12], [4]]
    eo = e0.clone().requires_grad_(True)
    loss = dec.forward_transducer(eo, elens, ys)
    loss.sum().mul(0.7).backward()                 # a non-unit upstream gradient, as rnnt_weight < 1 produces
    ours = {k: p.grad.clone() for k, p in dec.named_parameters() if p.grad is not None}
    ours_e = eo.grad.clone()
    for p in dec.parameters():
        p.grad = None
    # the same chain in plain torch
    U = 5
    ys_in = torch.full((B, U + 1), 3, dtype=torch.long)
    ys_out = torch.zeros(B, U, dtype=torch.int32)
    for b, y in enumerate(ys):
        ys_in[b, 0] = 2
        ys_in[b, 1:len(y) + 1] = torch.tensor(y)
        ys_out[b, :len(y)] = torch.tensor(y, dtype=torch.int32)
    er = e0.clone().requires_grad_(True)
    d = dec.embed(ys_in.cuda())
    for l in range(2):
        d, _ = dec.rnn[l](d)
        if dec.proj is not None:
            d = torch.relu(dec.proj[l](d))
    z = torch.tanh(dec.w_enc(er)[:, :, None] + dec.w_dec(d)[:, None])
    lp = torch.log_softmax(dec.output(z), -1)
    ref = torchaudio.functional.rnnt_loss(lp, ys_out.cuda(), elens.cuda(), torch.tensor([5, 3, 1], dtype=torch.int32).cuda(),
                                          blank=0, reduction="mean", fused_log_softmax=False)
    assert abs(loss.item() - ref.item()) <= (1e-3 if precision == "fp32" else 3e-2) * abs(ref.item()), (loss.item(), ref.item())
    ref.mul(0.7).backward()
    assert float((ours_e - er.grad).abs().max()) <= tol * float(er.grad.abs().max())
    gmax = max(float(p.grad.abs().max()) for p in dec.parameters() if p.grad is not None)
    bad = []
    for k, p in dec.named_parameters():
        if p.grad is None:
            continue
        assert k in ours, k
        e = float((ours[k] - p.grad).abs().max() / max(float(p.grad.abs().max()), 1e-3 * gmax))
        if not e <= tol:
            bad.append((k, e))
    assert not bad, (bad[:8], len(bad))


@pytest.mark.parametrize("name", ["rnnt_small.npz", "rnnt_mid.npz"])
@pytest.mark.parametrize("mode", ["fp32", "fp32_inplace", "bf16"])
def test_rnnt_grad_logits_matches_reference_fixture(name, mode):
    """One-pass d loss / d logits (log-softmax backward folded into the lattice gradient) against the fixture's
    `grad_logits` (torch autograd through log_softmax + torchaudio rnnt_loss, tests/golden/gen_golden_encoder.py::gen_rnnt)."""
    from conftest import load_golden
    from neural_sp_b200 import ops
    g = load_golden(name)
    dev = "cuda"
    logits = torch.from_numpy(g["logits"]).to(dev)
    lp = ops.softmax_rows(logits, log=True)
    ys, flens, ylens = (torch.from_numpy(g[k]).to(dev) for k in ("ys", "flens", "ylens"))
    loss, nll, _, ws = ops.rnnt_loss_fwd_bwd(lp, ys, flens, ylens, 0, need_grad=False, return_ws=True)
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    scale = torch.tensor(0.7, device=dev)
    ref = torch.from_numpy(g["grad_logits"]).to(dev) * 0.7
    if mode == "bf16":
        dz = ops.rnnt_grad_logits(lp, ws, nll, ys, flens, ylens, 0, scale, out_dtype=torch.bfloat16)
        tol = 1e-2
    else:
        dz = ops.rnnt_grad_logits(lp, ws, nll, ys, flens, ylens, 0, scale, inplace=(mode == "fp32_inplace"))
        tol = 2e-4
        assert (dz.data_ptr() == lp.data_ptr()) == (mode == "fp32_inplace")
    assert float((dz.float() - ref).abs().max()) <= tol * max(float(ref.abs().max()), 1e-3)
    # unit upstream gradient through the null pointer
    lp2 = ops.softmax_rows(logits, log=True)
    dz1 = ops.rnnt_grad_logits(lp2, ws, nll, ys, flens, ylens, 0)
    assert float((dz1 - ref / 0.7).abs().max()) <= 2e-4 * max(float(ref.abs().max()), 1e-3)
