"""Host logic of the Transformer / Conformer encoders -- offline AND chunk-by-chunk streaming -- pinned to the UNMODIFIED
reference on CPU (reference contract: test/encoders/test_transformer_encoder_streaming_chunkwise.py).

The library's ops are replaced by torch restatements of their contracts (tests/ops_doubles.py), the reference's weights are
loaded with a strict `load_state_dict`, and for every configuration three things are compared:
  1. offline forward: ours == reference (wiring, chunking, masks-as-kernel-parameters, length arithmetic);
  2. every streamed chunk: ours == the reference's streamed chunk (same inputs, `streaming=True`, lookback / lookahead),
     including the per-layer caches (`input_san`, `input_conv`) and their truncation;
  3. the reference's own contract: concatenated chunks == offline output (atol 1e-4).
What Python does with the kernels' results is therefore checked here; the kernels themselves (including attention with
`mlen > 0` cached keys) are checked on the GPU (tests/test_modules_gpu.py, tests/test_zz_streaming_gpu.py).
Needs /root/reference (build container only): skipped elsewhere."""
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


def make_args_transformer(**kw):
    a = dict(input_dim=80, enc_type='conv_transformer', n_heads=4, n_layers=3, n_layers_sub1=0, n_layers_sub2=0, d_model=8,
             d_ff=16, ffn_bottleneck_dim=0, ffn_activation='relu', pe_type='add', layer_norm_eps=1e-12, last_proj_dim=0,
             dropout_in=0.1, dropout=0.1, dropout_att=0.1, dropout_layer=0.1, subsample="1_1_1", subsample_type='max_pool',
             n_stacks=1, n_splices=1, frontend_conv=None, task_specific_layer=False, param_init='xavier_uniform',
             clamp_len=-1, lookahead="0", chunk_size_left="0", chunk_size_current="0", chunk_size_right="0",
             streaming_type='mask')
    a.update(kw)
    return a


def make_args_conformer(**kw):
    a = make_args_transformer(enc_type='conv_conformer', ffn_activation='swish', pe_type='relative')
    a.update(kernel_size=7, normalization='layer_norm')
    a.update(kw)
    return a


def make_args_conv(**kw):
    a = dict(input_dim=80, in_channel=1, channels="32_32", kernel_sizes="(3,3)_(3,3)", strides="(1,1)_(1,1)",
             poolings="(2,2)_(2,2)", dropout=0.1, normalization='', residual=False, bottleneck_dim=0, param_init=0.1)
    a.update(kw)
    return a


CASES = [
    # unidirectional, no CNN (cache grows by one chunk per call; causal mask with cached keys)
    ({'enc_type': 'uni_transformer', 'chunk_size_current': "1"}, None),
    ({'enc_type': 'uni_transformer', 'chunk_size_current': "4"}, None),
    ({'enc_type': 'uni_transformer', 'chunk_size_current': "4", 'pe_type': 'none'}, None),
    ({'enc_type': 'uni_transformer', 'chunk_size_current': "4", 'pe_type': 'relative_xl'}, None),
    ({'enc_type': 'uni_transformer', 'chunk_size_current': "4", 'lookahead': "1_0_1"}, None),
    ({'enc_type': 'uni_conformer', 'chunk_size_current': "1"}, None),
    ({'enc_type': 'uni_conformer', 'chunk_size_current': "4"}, None),
    ({'enc_type': 'uni_conformer', 'chunk_size_current': "4", 'pe_type': 'relative_xl'}, None),
    ({'enc_type': 'uni_conformer', 'chunk_size_current': "4", 'clamp_len': 5}, None),
    ({'enc_type': 'uni_conformer_v2', 'chunk_size_current': "1"}, None),
    ({'enc_type': 'uni_conformer_v2', 'chunk_size_current': "4"}, None),
    ({'enc_type': 'conformer_v2', 'streaming_type': 'reshape', 'chunk_size_left': "8", 'chunk_size_current': "16",
      'chunk_size_right': "8"}, None),
    ({'enc_type': 'conformer_v2', 'streaming_type': 'mask', 'chunk_size_left': "16", 'chunk_size_current': "8"}, None),
    ({'enc_type': 'conv_uni_conformer_v2', 'chunk_size_current': "16", 'subsample': "1_2_1"}, {}),
    ({'enc_type': 'conv_conformer_v2', 'streaming_type': 'mask', 'chunk_size_left': "16", 'chunk_size_current': "8"}, {}),
    # latency-controlled, no CNN
    ({'enc_type': 'transformer', 'streaming_type': 'reshape', 'chunk_size_left': "8", 'chunk_size_current': "16",
      'chunk_size_right': "8"}, None),
    ({'enc_type': 'transformer', 'streaming_type': 'mask', 'chunk_size_left': "16", 'chunk_size_current': "8"}, None),
    ({'enc_type': 'conformer', 'streaming_type': 'reshape', 'chunk_size_left': "8", 'chunk_size_current': "16",
      'chunk_size_right': "8"}, None),
    ({'enc_type': 'conformer', 'streaming_type': 'mask', 'chunk_size_left': "16", 'chunk_size_current': "8"}, None),
    # CNN front-ends: 1/2, 1/4 (lookback / lookahead trimming for the unidirectional ones), hierarchical 1/8
    ({'enc_type': 'conv', 'chunk_size_current': "2"},
     {'channels': "32", 'kernel_sizes': "(3,3)", 'strides': "(1,1)", 'poolings': "(2,2)"}),
    ({'enc_type': 'conv_transformer', 'streaming_type': 'mask', 'chunk_size_left': "16", 'chunk_size_current': "8"},
     {'channels': "32", 'kernel_sizes': "(3,3)", 'strides': "(1,1)", 'poolings': "(2,2)"}),
    ({'enc_type': 'conv', 'chunk_size_current': "8"}, {}),
    ({'enc_type': 'conv_uni_transformer', 'chunk_size_current': "8"}, {}),
    ({'enc_type': 'conv_uni_conformer', 'chunk_size_current': "8"}, {}),
    ({'enc_type': 'conv_transformer', 'streaming_type': 'mask', 'chunk_size_left': "16", 'chunk_size_current': "8"}, {}),
    ({'enc_type': 'conv_conformer', 'streaming_type': 'mask', 'chunk_size_left': "16", 'chunk_size_current': "8"}, {}),
    ({'enc_type': 'conv_transformer', 'streaming_type': 'reshape', 'chunk_size_left': "8", 'chunk_size_current': "16",
      'chunk_size_right': "8"}, {}),
    ({'enc_type': 'conv_conformer', 'streaming_type': 'reshape', 'chunk_size_left': "8", 'chunk_size_current': "16",
      'chunk_size_right': "8"}, {}),
    ({'enc_type': 'conv_conformer', 'streaming_type': 'mask', 'chunk_size_left': "16", 'chunk_size_current': "8",
      'subsample': "2_2_1"}, {'poolings': "(1,1)_(2,2)"}),
    # Conformer + "ConvSubsample": strided convolutions instead of pooling
    ({'enc_type': 'conv_conformer', 'streaming_type': 'mask', 'chunk_size_left': "16", 'chunk_size_current': "8"},
     {'channels': "32", 'kernel_sizes': "(3,3)", 'strides': "(2,2)", 'poolings': "(1,1)"}),
    ({'enc_type': 'conv_conformer', 'streaming_type': 'mask', 'chunk_size_left': "16", 'chunk_size_current': "8"},
     {'channels': "32_32", 'kernel_sizes': "(3,3)_(3,3)", 'strides': "(2,2)_(2,2)", 'poolings': "(1,1)_(1,1)"}),
    ({'enc_type': 'conv_conformer', 'streaming_type': 'mask', 'chunk_size_left': "16", 'chunk_size_current': "8",
      'subsample': "2_2_1"}, {'strides': "(1,1)_(2,2)", 'poolings': "(1,1)_(1,1)"}),
]


def _build_pair(args, args_conv):
    """(reference encoder, ours) with identical weights, both in eval mode, ours in parity (fp32) precision."""
    from oracle.ref_import import import_reference
    import_reference()
    from neural_sp_b200.encoders.conformer import ConformerEncoder
    from neural_sp_b200.encoders.conv import ConvEncoder
    from neural_sp_b200.encoders.transformer import TransformerEncoder
    is_conformer = 'conformer' in args['enc_type']
    a_ref, a_our = dict(args), dict(args)
    if args_conv is not None:
        ref_conv = importlib.import_module('neural_sp.models.seq2seq.encoders.conv')
        c = make_args_conv(**args_conv)
        c['bottleneck_dim'] = args['d_model']
        a_ref['frontend_conv'] = ref_conv.ConvEncoder(**c)
        a_our['frontend_conv'] = ConvEncoder(**c)
    if is_conformer:
        ref_mod = importlib.import_module('neural_sp.models.seq2seq.encoders.conformer')
        ref, ours = ref_mod.ConformerEncoder(**a_ref), ConformerEncoder(**a_our)
    else:
        ref_mod = importlib.import_module('neural_sp.models.seq2seq.encoders.transformer')
        ref, ours = ref_mod.TransformerEncoder(**a_ref), TransformerEncoder(**a_our)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.set_precision("fp32")
    return ref.eval(), ours.eval()


@pytest.mark.parametrize("ov, ov_conv", CASES)
def test_streaming_chunks_match_reference_and_offline(ov, ov_conv, monkeypatch):
    import ops_doubles
    ops_doubles.install(monkeypatch)
    torch.manual_seed(0)
    rng = np.random.RandomState(0)
    is_conformer = 'conformer' in ov['enc_type']
    args = make_args_conformer(**ov) if is_conformer else make_args_transformer(**ov)
    unidir = 'uni' in args['enc_type']
    N_l = max(0, int(args['chunk_size_left'])) // args['n_stacks']
    N_c = max(0, int(args['chunk_size_current'])) // args['n_stacks']
    N_r = max(0, int(args['chunk_size_right'])) // args['n_stacks']
    if unidir or args['enc_type'] == 'conv':
        args['chunk_size_left'] = args['chunk_size_current'] = args['chunk_size_right'] = "0"
    ref, ours = _build_pair(args, ov_conv)
    assert ours.cache_sizes == ref.cache_sizes
    if args['streaming_type'] == 'mask':
        N_l = 0                                     # previous chunks live in the encoder's cache
    factor = ref.subsampling_factor
    assert ours.subsampling_factor == factor
    conv_context = ref.conv.context_size if (ref.conv is not None and not ref.lc_bidir) else 0
    if ref.conv is not None:
        assert ours.conv.context_size == ref.conv.context_size
    bs, atol = 1, 1e-4

    for xmax_orig in (132, 141, 150):
        xs = rng.randn(bs, xmax_orig, args['input_dim']).astype(np.float32)
        if ref.streaming_type == 'mask' and ref.conv is not None and xmax_orig % N_c != 0:
            xs = np.concatenate([xs, np.zeros((bs, N_c - xmax_orig % N_c, args['input_dim']), np.float32)], axis=1)
        xs_pad = torch.from_numpy(xs)
        xlens = torch.IntTensor([xs.shape[1]] * bs)
        xmax = xs.shape[1]

        # 1. offline
        ref.reset_cache(), ours.reset_cache()
        with torch.no_grad():
            r_all = ref(xs_pad.clone(), xlens.clone(), task='all')['ys']
        o_all = ours(xs_pad.clone(), xlens.clone(), task='all')['ys']
        assert torch.equal(r_all['xlens'], o_all['xlens'])
        assert r_all['xs'].shape == o_all['xs'].shape
        assert torch.allclose(r_all['xs'], o_all['xs'], atol=atol), (r_all['xs'] - o_all['xs']).abs().max()
        eout_all = o_all['xs']

        # 2./3. chunk by chunk (loop structure of the reference's streaming test)
        n_chunks = math.ceil(xmax / N_c)
        j = j_out = 0
        cat, elens_cat = [], 0
        ref.reset_cache(), ours.reset_cache()
        for _ in range(n_chunks):
            start, end = j - N_l - conv_context, (j + N_c + N_r) + conv_context
            chunk = xs_pad[:, max(0, start):end]
            if ref.streaming_type == 'reshape':
                xlens_chunk = torch.IntTensor([max(factor, min(xmax - j, N_c))] * bs)
                if start < 0:
                    chunk = torch.cat([chunk.new_zeros(bs, -start, chunk.size(2)), chunk], dim=1)
                if end >= xmax:
                    chunk = torch.cat([chunk, chunk.new_zeros(bs, end - xmax, chunk.size(2))], dim=1)
            else:
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
            if args['enc_type'] != 'conv':
                assert ours.offset == ref.offset
                for lth in range(ours.n_layers):        # caches: same keys, same number of frames, same content
                    rc, oc = ref.cache[lth], ours.cache[lth]
                    assert (rc is None) == (oc is None)
                    if rc is not None:
                        assert sorted(rc) == sorted(k for k in oc if not k.startswith('_'))     # '_kv': ours only
                        n_frames = rc['input_san'].size(1)
                        assert all(t.size(1) == n_frames for t in oc['_kv'])
                        for key in rc:
                            assert rc[key].shape == oc[key].shape, (lth, key, rc[key].shape, oc[key].shape)
                            assert torch.allclose(rc[key], oc[key].float(), atol=atol)

            eout_all_i = eout_all[:, j_out:]
            if lookahead or conv_context == 0 or not unidir:
                eout_all_i = eout_all_i[:, :(N_c // factor)]
            if eout_all_i.size(1) == 0:
                break
            eout_chunk, elens_chunk = o_ck['xs'], o_ck['xlens'].clone()
            diff = eout_chunk.size(1) - eout_all_i.size(1)
            cat.append(eout_chunk[:, :eout_all_i.size(1)])
            elens_cat = elens_cat + (elens_chunk - diff)
            j += N_c
            j_out += N_c // factor
            if j > xmax:
                break
            if not lookahead and conv_context > 0 and unidir:
                break
        if sum(int(v) for v in args['lookahead'].split('_')) > 0:
            continue            # look-ahead frames beyond the chunk are not available when streaming: 1. and 2. only
        cat = torch.cat(cat, dim=1)
        assert cat.shape == eout_all.shape
        assert torch.allclose(eout_all, cat, atol=atol), (eout_all - cat).abs().max()
        assert torch.equal(o_all['xlens'], elens_cat)
