"""RNN-T training path on CPU (host wiring): prediction network (embedding, LSTM layer nodes, projections), joint network and
loss as autograd nodes (neural_sp_b200/autograd.py _RnntJointLossFn) against the UNMODIFIED reference `RNNTransducer`
(decoders/rnn_transducer.py:174-311) with identical weights: loss value, d loss / d encoder output and every parameter
gradient.  The reference's loss library is not installable offline (warp_rnnt / warprnnt_pytorch); a stand-in
module `warprnnt_pytorch` backed by torchaudio.functional.rnnt_loss (the oracle of tests/golden/rnnt_*.npz) is injected.
Needs /root/reference (build container only): skipped elsewhere."""
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/neural_sp"), reason="reference tree not available")


def _fake_warprnnt():
    import torchaudio
    mod = types.ModuleType("warprnnt_pytorch")

    class RNNTLoss(torch.nn.Module):
        def forward(self, log_probs, labels, flens, ylens):          # already normalised by the batch size (:257)
            return torchaudio.functional.rnnt_loss(log_probs, labels.int(), flens.int(), ylens.int(), blank=0,
                                                   reduction='mean', fused_log_softmax=False)
    mod.RNNTLoss = RNNTLoss
    return mod


@pytest.mark.parametrize("n_projs,ctc_weight", [(0, 0.0), (12, 0.3)])
def test_rnnt_training_matches_reference(n_projs, ctc_weight, monkeypatch):
    import ops_doubles
    from oracle.ref_import import import_reference
    import_reference()
    import neural_sp.models.seq2seq.decoders.rnn_transducer as ref_mod
    from neural_sp_b200.decoders.rnn_transducer import RNNTransducer
    ops_doubles.install_training(monkeypatch)
    monkeypatch.setitem(sys.modules, "warprnnt_pytorch", _fake_warprnnt())
    torch.manual_seed(0)
    sym = {'eos': 2, 'unk': 1, 'pad': 3, 'blank': 0}
    kw = dict(special_symbols=sym, enc_n_units=24, n_units=16, n_projs=n_projs, n_layers=2, bottleneck_dim=20, emb_dim=8,
              vocab=30, dropout=0.0, dropout_emb=0.0, ctc_weight=ctc_weight, ctc_lsm_prob=0.0, ctc_fc_list="", external_lm=None,
              global_weight=1.0, mtl_per_batch=False, param_init=0.1)
    ref = ref_mod.RNNTransducer(**kw).train()
    ours = RNNTransducer(**kw)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.set_precision("fp32")
    ours.train()
    B, T = 3, 14
    rng = np.random.RandomState(1)
    e0 = torch.from_numpy(rng.randn(B, T, 24).astype(np.float32))
    elens = torch.IntTensor([14, 11, 9])
    ys = [[5, 6, 7, 8], [9, 10], [4]]
    e_ref = e0.clone().requires_grad_(True)
    e_our = e0.clone().requires_grad_(True)
    loss_r, obs_r = ref(e_ref, elens.clone(), ys, task='all')
    loss_o, obs_o = ours(e_our, elens.clone(), ys, task='all')
    assert loss_o.shape == loss_r.shape == (1,)
    assert abs(float(loss_o) - float(loss_r)) <= 1e-4 * abs(float(loss_r)), (float(loss_o), float(loss_r))
    assert abs(obs_o['loss_transducer'] - obs_r['loss_transducer']) <= 1e-4 * abs(obs_r['loss_transducer'])
    loss_r.sum().backward()
    loss_o.sum().backward()
    assert float((e_our.grad - e_ref.grad).abs().max()) <= 1e-4 * float(e_ref.grad.abs().max())
    ref_g = dict(ref.named_parameters())
    for k, p in ours.named_parameters():
        g = ref_g[k].grad
        if g is None:
            continue
        assert p.grad is not None, k
        err = float((p.grad - g).abs().max() / g.abs().max().clamp_min(1e-12))
        assert err <= 2e-4, (k, err)
    # eval mode / no grad: the inference kernels' path (prediction network on the LSTM kernel, materialised log_probs)
    ref.eval(), ours.eval()
    with torch.no_grad():
        l_r = ref.forward_transducer(e0.clone(), elens.clone(), ys)
        l_o = ours.forward_transducer(e0.clone(), elens.clone(), ys)
    assert abs(float(l_o) - float(l_r)) <= 1e-4 * abs(float(l_r)), (float(l_o), float(l_r))


@pytest.mark.parametrize("ov", [{'n_layers': 1}, {'n_layers': 2}, {'n_projs': 8}, {'ctc_weight': 0.5}, {'ctc_weight': 1.0},
                                {'ctc_weight': 1.0, 'ctc_lsm_prob': 0.0}])
def test_reference_rnnt_test_matrix(ov, monkeypatch):
    """The reference's own decoder test matrix (test/decoders/test_rnn_transducer_decoder.py::test_forward: numpy label arrays that
    may contain the blank id, auxiliary CTC with two fc layers and label smoothing, pure-CTC weighting), dropouts set to 0:
    total loss, observation dict and all gradients against the unmodified reference."""
    import ops_doubles
    from oracle.ref_import import import_reference
    import_reference()
    import neural_sp.models.seq2seq.decoders.rnn_transducer as ref_mod
    from neural_sp_b200.decoders.rnn_transducer import RNNTransducer
    ops_doubles.install_training(monkeypatch)
    monkeypatch.setitem(sys.modules, "warprnnt_pytorch", _fake_warprnnt())
    torch.manual_seed(0)
    kw = dict(special_symbols={'blank': 0, 'unk': 1, 'eos': 2, 'pad': 3}, enc_n_units=16, n_units=16, n_projs=0, n_layers=2,
              bottleneck_dim=8, emb_dim=8, vocab=10, dropout=0.0, dropout_emb=0.0, ctc_weight=0.1, ctc_lsm_prob=0.1,
              ctc_fc_list='16_16', external_lm=None, global_weight=1.0, mtl_per_batch=False, param_init=0.1)
    kw.update(ov)
    ref = ref_mod.RNNTransducer(**kw).train()
    ours = RNNTransducer(**kw)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.set_precision("fp32")
    ours.train()
    rng = np.random.RandomState(2)
    e0 = torch.from_numpy(rng.randn(4, 40, 16).astype(np.float32))
    elens = torch.IntTensor([40] * 4)
    ys = [rng.randint(0, 10, n).astype(np.int32) for n in (4, 5, 3, 7)]
    e_r, e_o = e0.clone().requires_grad_(True), e0.clone().requires_grad_(True)
    loss_r, obs_r = ref(e_r, elens.clone(), ys, task='all')
    loss_o, obs_o = ours(e_o, elens.clone(), ys, task='all')
    assert loss_o.dim() == 1 and loss_o.size(0) == 1 and isinstance(obs_o, dict)
    assert abs(float(loss_o.detach()) - float(loss_r.detach())) <= 1e-4 * abs(float(loss_r.detach()))
    for k, v in obs_r.items():
        if isinstance(v, float):
            assert abs(obs_o[k] - v) <= 1e-4 * max(1.0, abs(v)), k
    loss_r.sum().backward()
    loss_o.sum().backward()
    assert float((e_o.grad - e_r.grad).abs().max()) <= 2e-4 * float(e_r.grad.abs().max())
    rg = dict(ref.named_parameters())
    for k, p in ours.named_parameters():
        g = rg[k].grad
        if g is None:
            continue
        err = float((p.grad - g).abs().max() / g.abs().max().clamp_min(1e-12))
        assert err <= 5e-4, (k, err)
