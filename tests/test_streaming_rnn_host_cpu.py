"""Host logic of the (CNN-)LSTM encoders -- offline, streaming state carry-over and the latency-controlled BLSTM chunk loop --
pinned to the UNMODIFIED reference on CPU (reference contract: test/encoders/test_rnn_encoder_streaming_chunkwise.py).
Same method as tests/test_streaming_host_cpu.py: ops replaced by their torch restatements (tests/ops_doubles.py), the
reference's weights loaded with a strict `load_state_dict`; compared: offline output, every streamed chunk (and the carried
(h, c) states), and the reference's own contract chunks == offline."""
import importlib
import math
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/neural_sp"), reason="reference tree not available")


def make_args(**kw):
    a = dict(input_dim=80, enc_type='blstm', n_units=16, n_projs=0, last_proj_dim=0, n_layers=2, n_layers_sub1=0,
             n_layers_sub2=0, dropout_in=0.1, dropout=0.1, subsample="1_1", subsample_type='drop', n_stacks=1, n_splices=1,
             frontend_conv=None, bidir_sum_fwd_bwd=False, task_specific_layer=False, param_init=0.1,
             chunk_size_current="0", chunk_size_right="0", cnn_lookahead=True, rsp_prob=0)
    a.update(kw)
    return a


def make_args_conv(**kw):
    a = dict(input_dim=80, in_channel=1, channels="32_32", kernel_sizes="(3,3)_(3,3)", strides="(1,1)_(1,1)",
             poolings="(2,2)_(2,2)", dropout=0.1, normalization='', residual=False, bottleneck_dim=0, param_init=0.1)
    a.update(kw)
    return a


ONE_BLOCK = {'channels': "32", 'kernel_sizes': "(3,3)", 'strides': "(1,1)", 'poolings': "(2,2)"}
CASES = [
    ({'enc_type': 'blstm', 'chunk_size_current': "20", 'chunk_size_right': "20"}, None),
    ({'enc_type': 'blstm', 'chunk_size_current': "32", 'chunk_size_right': "16"}, None),
    ({'enc_type': 'blstm', 'chunk_size_current': "32", 'chunk_size_right': "16", 'bidir_sum_fwd_bwd': True, 'n_projs': 8}, None),
    ({'enc_type': 'lstm', 'chunk_size_current': "1"}, None),
    ({'enc_type': 'lstm', 'chunk_size_current': "40"}, None),
    ({'enc_type': 'lstm', 'chunk_size_current': "40", 'n_projs': 8, 'last_proj_dim': 12}, None),
    ({'enc_type': 'conv_blstm', 'chunk_size_current': "32", 'chunk_size_right': "16"}, ONE_BLOCK),
    ({'enc_type': 'conv_blstm', 'n_layers': 3, 'subsample': '1_2_1', 'chunk_size_current': "32", 'chunk_size_right': "16"},
     ONE_BLOCK),
    ({'enc_type': 'conv_lstm', 'chunk_size_current': "8"}, {}),
    ({'enc_type': 'conv_lstm', 'chunk_size_current': "40"}, {}),
    ({'enc_type': 'conv_blstm', 'chunk_size_current': "32", 'chunk_size_right': "16"}, {}),
    ({'enc_type': 'conv_blstm', 'cnn_lookahead': False, 'chunk_size_current': "32", 'chunk_size_right': "16"}, {}),
]


@pytest.mark.parametrize("ov, ov_conv", CASES)
def test_rnn_streaming_chunks_match_reference_and_offline(ov, ov_conv, monkeypatch):
    import ops_doubles
    ops_doubles.install(monkeypatch)
    from oracle.ref_import import import_reference
    import_reference()
    from neural_sp_b200.encoders.conv import ConvEncoder
    from neural_sp_b200.encoders.rnn import RNNEncoder
    torch.manual_seed(0)
    rng = np.random.RandomState(0)
    args = make_args(**ov)
    unidir = args['enc_type'] in ['conv_lstm', 'lstm']
    N_c = int(args['chunk_size_current']) // args['n_stacks']
    N_r = int(args['chunk_size_right']) // args['n_stacks']
    if unidir:
        args['chunk_size_current'] = args['chunk_size_right'] = "0"
    a_ref, a_our = dict(args), dict(args)
    if ov_conv is not None:
        c = make_args_conv(**ov_conv)
        a_ref['frontend_conv'] = importlib.import_module('neural_sp.models.seq2seq.encoders.conv').ConvEncoder(**c)
        a_our['frontend_conv'] = ConvEncoder(**c)
    ref = importlib.import_module('neural_sp.models.seq2seq.encoders.rnn').RNNEncoder(**a_ref).eval()
    ours = RNNEncoder(**a_our)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.set_precision("fp32")
    ours.eval()
    factor = ref.subsampling_factor
    assert ours.subsampling_factor == factor and ours.output_dim == ref.output_dim
    conv_context = ref.conv.context_size if ref.conv is not None else 0
    bs, atol = 1, 1e-5

    for xmax in (160, 175, 187):
        xs_pad = torch.from_numpy(rng.randn(bs, xmax, 80).astype(np.float32))
        xlens = torch.IntTensor([xmax] * bs)
        ref.reset_cache(), ours.reset_cache()
        with torch.no_grad():
            r_all = ref(xs_pad.clone(), xlens.clone(), task='all')['ys']
        o_all = ours(xs_pad.clone(), xlens.clone(), task='all')['ys']
        assert torch.equal(r_all['xlens'], o_all['xlens'])
        assert r_all['xs'].shape == o_all['xs'].shape
        assert torch.allclose(r_all['xs'], o_all['xs'], atol=atol), (r_all['xs'] - o_all['xs']).abs().max()
        eout_all = o_all['xs']

        j = j_out = 0
        cat, elens_cat = [], 0
        ref.reset_cache(), ours.reset_cache()
        for _ in range(math.ceil(xmax / N_c)):
            start, end = j - conv_context, (j + N_c + N_r) + conv_context
            chunk = xs_pad[:, max(0, start):end]
            xlens_chunk = torch.IntTensor([max(factor, chunk.size(1))] * bs)
            lookback = start >= 0 and conv_context > 0
            lookahead = end < xmax and conv_context > 0
            with torch.no_grad():
                r_ck = ref(chunk.clone(), xlens_chunk.clone(), task='all', streaming=True, lookback=lookback,
                           lookahead=lookahead)['ys']
            o_ck = ours(chunk.clone(), xlens_chunk.clone(), task='all', streaming=True, lookback=lookback,
                        lookahead=lookahead)['ys']
            assert torch.equal(r_ck['xlens'], o_ck['xlens']), (r_ck['xlens'], o_ck['xlens'])
            assert r_ck['xs'].shape == o_ck['xs'].shape
            assert torch.allclose(r_ck['xs'], o_ck['xs'], atol=atol), (r_ck['xs'] - o_ck['xs']).abs().max()
            for lth in range(ours.n_layers):            # carried (h_n, c_n): same layout and values as nn.LSTM's
                for rs, os_ in zip(ref.hx_fwd[lth], ours.hx_fwd[lth]):
                    assert rs.shape == os_.shape and torch.allclose(rs, os_, atol=atol)

            eout_all_i = eout_all[:, j_out:]
            if lookahead or conv_context == 0 or not unidir:
                eout_all_i = eout_all_i[:, :(N_c // factor)]
            if eout_all_i.size(1) == 0:
                break
            diff = o_ck['xs'].size(1) - eout_all_i.size(1)
            cat.append(o_ck['xs'][:, :eout_all_i.size(1)])
            elens_cat = elens_cat + (o_ck['xlens'].clone() - diff)
            j += N_c
            j_out += N_c // factor
            if j > xmax:
                break
            if not lookahead and conv_context > 0 and unidir:
                break
        cat = torch.cat(cat, dim=1)
        assert cat.shape == eout_all.shape
        assert torch.allclose(eout_all, cat, atol=atol), (eout_all - cat).abs().max()
        assert torch.equal(o_all['xlens'], elens_cat)
