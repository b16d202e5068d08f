"""Drop-in under the UNCHANGED reference facade (INTEGRATION.md level 1), CPU only.

Builds the reference's `Speech2Text` twice from the same argument namespace -- once stock, once with
`build_encoder` / `CTC` resolved to neural_sp_b200's classes -- and checks that the facade cannot tell the difference at
construction level: same parameter names and shapes everywhere, strict `load_state_dict` of the stock weights, same
encoder properties the facade and the decoders read.  Needs the reference tree (/root/reference, present in the build
container only): skipped elsewhere.  Forward/backward through the facade needs a GPU and is covered by the module tests."""
import argparse
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/neural_sp"), reason="reference tree not available")


def make_args(**ov):
    a = dict(input_type='speech', input_dim=80, enc_type='conv_conformer', dec_type='lstm', enc_n_layers=3, enc_n_layers_sub1=0,
             enc_n_layers_sub2=0, subsample="1_2_1", subsample_type='max_pool', vocab=40, vocab_sub1=-1, vocab_sub2=-1,
             total_weight=1.0, sub1_weight=0.0, sub2_weight=0.0, mtl_per_batch=False, task_specific_layer=False,
             ctc_weight=1.0, ctc_weight_sub1=0.0, ctc_weight_sub2=0.0, bwd_weight=0.0, mbr_training=False,
             input_noise_std=0.0, n_stacks=1, n_skips=1, n_splices=1, weight_noise_std=0.0, n_freq_masks=0, n_time_masks=0,
             freq_width=27, time_width=100, time_width_upper=1.0, adaptive_number_ratio=0.0, adaptive_size_ratio=0.0,
             max_n_time_masks=20, sequence_summary_network=False, freeze_encoder=False, external_lm=None, lm_fusion='',
             param_init=0.1, emb_dim=32, dropout_emb=0.0,
             conv_in_channel=1, conv_channels="32_32", conv_kernel_sizes="(3,3)_(3,3)", conv_strides="(1,1)_(1,1)",
             conv_poolings="(1,1)_(2,2)", conv_normalization='', conv_bottleneck_dim=0,
             transformer_enc_n_heads=4, transformer_enc_d_model=64, transformer_enc_d_ff=128, transformer_enc_pe_type='relative',
             transformer_enc_clamp_len=10, transformer_enc_lookaheads="0_0_0", transformer_ffn_bottleneck_dim=0,
             transformer_ffn_activation='swish', transformer_layer_norm_eps=1e-12, transformer_param_init='xavier_uniform',
             transformer_dec_d_model=64, conformer_kernel_size=7, conformer_normalization='layer_norm',
             dropout_in=0.0, dropout_enc=0.0, dropout_att=0.0, dropout_enc_layer=0.0,
             lc_chunk_size_left="0", lc_chunk_size_current="0", lc_chunk_size_right="0", lc_type='reshape',
             enc_n_units=32, enc_n_projs=0, bidirectional_sum_fwd_bwd=False, cnn_lookahead=True, rsp_prob_enc=0.0,
             dec_n_units=32, dec_n_projs=0, dec_n_layers=1, dec_bottleneck_dim=32, tie_embedding=False, attn_type='location',
             attn_dim=32, attn_sharpening_factor=1.0, attn_sigmoid=False, attn_conv_n_channels=10, attn_conv_width=201,
             attn_n_heads=1, dropout_dec=0.0, lsm_prob=0.0, ss_prob=0.0, ctc_lsm_prob=0.1, ctc_fc_list="32", mbr_ce_weight=0.01,
             lm_init=None, mocha_chunk_size=1, mocha_n_heads_mono=1, mocha_init_r=-4, mocha_eps=1e-6, mocha_std=1.0,
             mocha_no_denominator=False, mocha_1dconv=False, mocha_decot_lookahead=0, mocha_quantity_loss_weight=0.0,
             mocha_latency_metric='', mocha_latency_loss_weight=0.0, mocha_stableemit_weight=0.0, gmm_attn_n_mixtures=1,
             replace_sos=False, distillation_weight=0.0, discourse_aware=False)
    a.update(ov)
    return argparse.Namespace(**a)


@pytest.mark.parametrize("ov", [dict(), dict(enc_type='conv_transformer', transformer_enc_pe_type='relative_xl',
                                             transformer_ffn_activation='relu'),
                                dict(enc_type='blstm', subsample="1_1_1")])
def test_speech2text_accepts_b200_modules(ov, monkeypatch):
    from oracle.ref_import import import_reference
    import_reference()
    import neural_sp.models.seq2seq.decoders.ctc as ref_ctc
    import neural_sp.models.seq2seq.decoders.las as ref_las
    import neural_sp.models.seq2seq.speech2text as ref_s2t
    from neural_sp_b200.decoders.ctc import CTC as B200CTC
    from neural_sp_b200.encoders.build import build_encoder as b200_build_encoder

    import torch
    torch.manual_seed(0)
    stock = ref_s2t.Speech2Text(make_args(**ov))
    monkeypatch.setattr(ref_s2t, "build_encoder", b200_build_encoder)
    monkeypatch.setattr(ref_las, "CTC", B200CTC)
    monkeypatch.setattr(ref_ctc, "CTC", B200CTC)
    torch.manual_seed(0)
    ours = ref_s2t.Speech2Text(make_args(**ov))
    # same seed -> bit-identical fresh weights (every initialiser of the mirrors draws in the reference's order)
    diff = [k for k, v in stock.state_dict().items() if not torch.equal(v, ours.state_dict()[k])]
    assert not diff, diff[:8]

    assert type(ours.enc).__module__.startswith("neural_sp_b200.") and type(ours.dec_fwd.ctc).__module__.startswith("neural_sp_b200.")
    s_ref = {k: tuple(v.shape) for k, v in stock.state_dict().items()}
    s_our = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    assert s_our == s_ref, set(s_our) ^ set(s_ref)
    ours.load_state_dict(stock.state_dict(), strict=True)
    for prop in ("output_dim", "subsampling_factor", "enc_type"):
        assert getattr(ours.enc, prop) == getattr(stock.enc, prop), prop
    assert ours.dec_fwd.ctc.lsm_prob == stock.dec_fwd.ctc.lsm_prob


def test_speech2text_accepts_b200_rnn_transducer(monkeypatch):
    """RNN-T decoder (conv + LSTM encoder, transducer + auxiliary CTC): the factory imports the class at call time, so the
    alias is set on the module (decoders/build.py:66-85)."""
    from oracle.ref_import import import_reference
    import_reference()
    import neural_sp.models.seq2seq.decoders.rnn_transducer as ref_rnnt
    import neural_sp.models.seq2seq.speech2text as ref_s2t
    from neural_sp_b200.decoders.rnn_transducer import RNNTransducer as B200RNNT
    from neural_sp_b200.encoders.build import build_encoder as b200_build_encoder

    ov = dict(dec_type='lstm_transducer', enc_type='conv_lstm', subsample="1_1_1", ctc_weight=0.3, conv_poolings="(2,2)_(2,2)")
    import torch
    torch.manual_seed(0)
    stock = ref_s2t.Speech2Text(make_args(**ov))
    monkeypatch.setattr(ref_s2t, "build_encoder", b200_build_encoder)
    monkeypatch.setattr(ref_rnnt, "RNNTransducer", B200RNNT)
    torch.manual_seed(0)
    ours = ref_s2t.Speech2Text(make_args(**ov))
    diff = [k for k, v in stock.state_dict().items() if not torch.equal(v, ours.state_dict()[k])]
    assert not diff, diff[:8]
    assert type(ours.dec_fwd).__module__.startswith("neural_sp_b200.") and type(ours.enc).__module__.startswith("neural_sp_b200.")
    s_ref = {k: tuple(v.shape) for k, v in stock.state_dict().items()}
    s_our = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    assert s_our == s_ref, set(s_our) ^ set(s_ref)
    ours.load_state_dict(stock.state_dict(), strict=True)


def _fake_warprnnt():
    import types
    import torch
    import torchaudio
    mod = types.ModuleType("warprnnt_pytorch")

    class RNNTLoss(torch.nn.Module):
        def forward(self, log_probs, labels, flens, ylens):
            return torchaudio.functional.rnnt_loss(log_probs, labels.int(), flens.int(), ylens.int(), blank=0,
                                                   reduction='mean', fused_log_softmax=False)
    mod.RNNTLoss = RNNTLoss
    return mod


@pytest.mark.parametrize("ov", [
    dict(),                                                                                      # conv + Conformer + CTC
    dict(enc_type='conv_transformer', transformer_enc_pe_type='add', transformer_ffn_activation='relu'),
    dict(enc_type='conv_blstm', subsample="1_2_1", enc_n_projs=16),                              # conv + BLSTM + CTC
    dict(dec_type='lstm_transducer', enc_type='conv_lstm', subsample="1_1_1", ctc_weight=0.3, conv_poolings="(2,2)_(2,2)"),
    dict(n_freq_masks=2, n_time_masks=2, freq_width=13, time_width=20),                          # SpecAugment in encode()
])
def test_training_step_through_the_unchanged_facade(ov, monkeypatch):
    """End to end under the UNCHANGED reference `Speech2Text.forward` (speech2text.py:206-345): same batch, same weights,
    stock modules vs neural_sp_b200's -- total loss, the observation dict and every parameter gradient.  CPU: the library's
    ops are replaced by their torch restatements (tests/ops_doubles.py), so this pins the drop-in boundary (argument
    passing, return structures, autograd connectivity of encoder -> decoder), not the kernels."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import numpy as np
    import torch
    import ops_doubles
    from oracle.ref_import import import_reference
    import_reference()
    import neural_sp.models.seq2seq.decoders.ctc as ref_ctc
    import neural_sp.models.seq2seq.decoders.las as ref_las
    import neural_sp.models.seq2seq.decoders.rnn_transducer as ref_rnnt
    import neural_sp.models.seq2seq.speech2text as ref_s2t
    from neural_sp_b200.decoders.ctc import CTC as B200CTC
    from neural_sp_b200.decoders.rnn_transducer import RNNTransducer as B200RNNT
    from neural_sp_b200.encoders.build import build_encoder as b200_build_encoder
    monkeypatch.setitem(sys.modules, "warprnnt_pytorch", _fake_warprnnt())
    torch.manual_seed(0)
    stock = ref_s2t.Speech2Text(make_args(**ov))
    monkeypatch.setattr(ref_s2t, "build_encoder", b200_build_encoder)
    monkeypatch.setattr(ref_las, "CTC", B200CTC)
    monkeypatch.setattr(ref_ctc, "CTC", B200CTC)
    monkeypatch.setattr(ref_rnnt, "RNNTransducer", B200RNNT)
    # the facade's own `isinstance(self.dec_fwd, RNNT)` (:300, :337, :629); a tuple because this test runs BOTH models
    monkeypatch.setattr(ref_s2t, "RNNT", (ref_s2t.RNNT, B200RNNT))
    from neural_sp_b200.frontends.spec_augment import SpecAugment as B200SpecAugment
    monkeypatch.setattr(ref_s2t, "SpecAugment", B200SpecAugment)
    torch.manual_seed(0)
    ours = ref_s2t.Speech2Text(make_args(**ov))
    ours.load_state_dict(stock.state_dict(), strict=True)
    for m in ours.modules():
        m.precision = "fp32"
    if ov.get("n_freq_masks", 0) > 0:
        assert type(ours.specaug).__module__.startswith("neural_sp_b200.") and type(stock.specaug).__module__.startswith("neural_sp.")
    ops_doubles.install_training(monkeypatch)
    rng = np.random.RandomState(0)
    batch = {'xs': [rng.randn(n, 80).astype(np.float32) for n in (64, 53, 40)],
             'ys': [[5, 6, 7, 8, 9], [10, 11, 12], [13, 4]], 'ys_sub1': None, 'ys_sub2': None, 'trigger_points': None,
             'xlens': [64, 53, 40], 'utt_ids': ['a', 'b', 'c'], 'speakers': ['s'] * 3, 'sessions': ['x'] * 3, 'text': [''] * 3,
             'feat_path': [''] * 3, 'ylens': [5, 3, 2]}
    stock.train(), ours.train()
    np.random.seed(11)                       # SpecAugment draws from numpy's global generator (same order on both sides)
    loss_s, obs_s = stock(batch, task='all')
    np.random.seed(11)
    loss_o, obs_o = ours(batch, task='all')
    assert loss_o.shape == loss_s.shape
    assert abs(float(loss_o.detach()) - float(loss_s.detach())) <= 1e-4 * abs(float(loss_s.detach()))
    for k, v in obs_s.items():
        if isinstance(v, float):
            assert abs(obs_o[k] - v) <= 1e-4 * max(1.0, abs(v)), (k, obs_o[k], v)
    loss_s.sum().backward()
    loss_o.sum().backward()
    gs = dict(stock.named_parameters())
    gmax = max(float(p.grad.abs().max()) for p in gs.values() if p.grad is not None)
    bad = []
    for k, p in ours.named_parameters():
        g = gs[k].grad
        if g is None:
            continue
        assert p.grad is not None, k
        e = float((p.grad - g).abs().max() / max(float(g.abs().max()), 1e-3 * gmax))
        if not e <= 1e-3:
            bad.append((k, e))
    assert not bad, (bad[:8], len(bad))


@pytest.mark.parametrize("ov", [
    dict(enc_type='conv_uni_conformer', conv_poolings="(2,2)_(2,2)", subsample="1_1_1"),                 # unidirectional + CNN context
    dict(enc_type='conv_transformer', transformer_enc_pe_type='relative_xl', transformer_ffn_activation='relu',
         conv_poolings="(2,2)_(2,2)", subsample="1_1_1", lc_chunk_size_left="32", lc_chunk_size_current="32",
         lc_chunk_size_right="16", lc_type='reshape'),                                                   # LC Transformer, windows
    dict(enc_type='conv_conformer', conv_poolings="(2,2)_(2,2)", subsample="1_1_1", lc_chunk_size_left="32",
         lc_chunk_size_current="16", lc_chunk_size_right="0", lc_type='mask'),                           # LC Conformer, chunk mask
    dict(enc_type='conv_lstm', conv_poolings="(2,2)_(2,2)", subsample="1_1_1"),                          # LSTM state carry-over
    dict(enc_type='conv_blstm', conv_poolings="(2,2)_(2,2)", subsample="1_1_1", lc_chunk_size_left="32",
         lc_chunk_size_right="16", bidirectional_sum_fwd_bwd=True),           # LC-BLSTM (the factory reads N_c from _left, build.py:146)
])
def test_streaming_encode_through_the_unchanged_facade(ov, monkeypatch):
    """`Speech2Text.encode_streaming` (speech2text.py:513-549) with the reference's own `Streaming` block slicer driving the
    encoder chunk by chunk: the stock encoder vs neural_sp_b200's under the same facade object -- the cached / carried state
    contract (`reset_cache`, `streaming=True`, CNN lookback / lookahead, `xlen_block`) as the facade uses it."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import numpy as np
    import torch
    import ops_doubles
    from oracle.ref_import import import_reference
    import_reference()
    import neural_sp.models.seq2seq.decoders.ctc as ref_ctc
    import neural_sp.models.seq2seq.decoders.las as ref_las
    import neural_sp.models.seq2seq.speech2text as ref_s2t
    from neural_sp_b200.decoders.ctc import CTC as B200CTC
    from neural_sp_b200.encoders.build import build_encoder as b200_build_encoder
    ov = dict(ov, ctc_weight=0.5)            # hybrid model: encode_streaming asserts an attention decoder (fwd_weight > 0)
    torch.manual_seed(0)
    stock = ref_s2t.Speech2Text(make_args(**ov)).eval()
    monkeypatch.setattr(ref_s2t, "build_encoder", b200_build_encoder)
    monkeypatch.setattr(ref_las, "CTC", B200CTC)
    monkeypatch.setattr(ref_ctc, "CTC", B200CTC)
    torch.manual_seed(0)
    ours = ref_s2t.Speech2Text(make_args(**ov))
    ours.load_state_dict(stock.state_dict(), strict=True)
    for m in ours.modules():
        m.precision = "fp32"
    ours.eval()
    ops_doubles.install(monkeypatch)
    params = dict(recog_block_sync_size=40, recog_ctc_vad=False, recog_ctc_vad_blank_threshold=40,
                  recog_ctc_vad_spike_threshold=0.1, recog_ctc_vad_n_accum_frames=4000)
    rng = np.random.RandomState(0)
    for T in (170, 203):
        x = rng.randn(T, 80).astype(np.float32)
        with torch.no_grad():
            e_s, l_s = stock.encode_streaming([x], params, task='ys')
            e_o, l_o = ours.encode_streaming([x], params, task='ys')
        assert torch.equal(l_s, l_o) and e_s.shape == e_o.shape
        assert float((e_s - e_o).abs().max()) <= 1e-4 * float(e_s.abs().max())


@pytest.mark.parametrize("beam", [1, 4])
def test_ctc_decode_through_the_unchanged_facade(beam, monkeypatch):
    """`Speech2Text.decode` (speech2text.py:709-767) on a CTC-weighted model: greedy (`recog_beam_width` 1) and the prefix beam
    search (ctc.py:256-363) -- the stock decoder vs neural_sp_b200's CTC under the same facade, same weights: identical N-best
    token ids."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import numpy as np
    import torch
    import ops_doubles
    from oracle.ref_import import import_reference
    import_reference()
    import neural_sp.models.seq2seq.decoders.ctc as ref_ctc
    import neural_sp.models.seq2seq.decoders.las as ref_las
    import neural_sp.models.seq2seq.speech2text as ref_s2t
    from neural_sp_b200 import ops
    from neural_sp_b200.decoders.ctc import CTC as B200CTC
    from neural_sp_b200.encoders.build import build_encoder as b200_build_encoder
    ov = dict(enc_type='conv_conformer', conv_poolings="(2,2)_(2,2)", subsample="1_1_1", ctc_weight=0.5)
    torch.manual_seed(0)
    stock = ref_s2t.Speech2Text(make_args(**ov))
    monkeypatch.setattr(ref_s2t, "build_encoder", b200_build_encoder)
    monkeypatch.setattr(ref_las, "CTC", B200CTC)
    monkeypatch.setattr(ref_ctc, "CTC", B200CTC)
    torch.manual_seed(0)
    ours = ref_s2t.Speech2Text(make_args(**ov))
    ours.load_state_dict(stock.state_dict(), strict=True)
    for m in ours.modules():
        m.precision = "fp32"
    ops_doubles.install(monkeypatch)
    monkeypatch.setattr(ops, "softmax_rows", ops_doubles.softmax_rows)

    def greedy_double(logits, elens, blank=0):          # nsp_ctc_greedy restated: best path, collapsed + blank-free, lengths
        B, T, _ = logits.shape
        best = logits.argmax(-1).int()
        hyp, lens, trig = torch.zeros(B, T, dtype=torch.int32), torch.zeros(B, dtype=torch.int32), torch.zeros(B, T, dtype=torch.int32)
        for b in range(B):
            prev, n = -1, 0
            for t in range(int(elens[b])):
                c = int(best[b, t])
                if c != prev and c != blank:
                    hyp[b, n], trig[b, n] = c, t
                    n += 1
                prev = c
            lens[b] = n
        return best, hyp, lens, trig
    monkeypatch.setattr(ops, "ctc_greedy", greedy_double)
    rng = np.random.RandomState(3)
    xs = [rng.randn(n, 80).astype(np.float32) * 2 for n in (72, 60)]
    params = {'recog_beam_width': beam, 'recog_ctc_weight': 1.0, 'recog_length_penalty': 0.1, 'recog_lm_weight': 0.0,
              'recog_lm_second_weight': 0.0, 'recog_lm_bwd_weight': 0.0, 'recog_lm_state_carry_over': False,
              'recog_softmax_smoothing': 1.0, 'recog_cache_embedding': False, 'recog_streaming_encoding': False,
              'recog_bwd_attention': False}
    hs, _ = stock.decode(xs, params, None, utt_ids=['u1', 'u2'], speakers=['a', 'b'])
    ho, _ = ours.decode(xs, params, None, utt_ids=['u1', 'u2'], speakers=['a', 'b'])
    assert type(ours.dec_fwd.ctc).__module__.startswith("neural_sp_b200.")
    assert len(hs) == len(ho) == 2
    for a, b in zip(hs, ho):
        assert [list(map(int, x)) for x in a] == [list(map(int, x)) for x in b]
