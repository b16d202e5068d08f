#!/usr/bin/env python3
"""Streaming (chunk-by-chunk) encoder fixtures from the UNMODIFIED reference, run on CPU in the build container:

    python tests/golden/gen_golden_streaming.py

Each stream_*.npz holds the reference module's state_dict (``sd.<key>``), one seeded utterance ``xs`` `[1, T, 80]`, the
constructor arguments (``cfg`` JSON), the reference's OFFLINE output ``ys`` and the chunk schedule of the reference's own
streaming test (test/encoders/test_transformer_encoder_streaming_chunkwise.py:215-283): per chunk the input slice
``[start, end)`` with its zero padding, ``xlens``, the lookback / lookahead flags (``sched`` int32 `[n, 7]`:
start, end, pad_left, pad_right, xlen, lookback, lookahead) and the reference's streamed output ``ck.<i>`` with
``ck_lens``."""
import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_import import import_reference  # noqa: E402

import_reference()

BASE = dict(input_dim=80, enc_type='conv_conformer', n_heads=4, kernel_size=7, normalization='layer_norm',
            n_layers=3, n_layers_sub1=0, n_layers_sub2=0, d_model=32, d_ff=64, ffn_bottleneck_dim=0,
            ffn_activation='swish', pe_type='relative', layer_norm_eps=1e-12, last_proj_dim=0, dropout_in=0.0,
            dropout=0.0, dropout_att=0.0, dropout_layer=0.0, subsample="1_1_1", subsample_type='max_pool',
            n_stacks=1, n_splices=1, frontend_conv=None, task_specific_layer=False, param_init='xavier_uniform',
            clamp_len=10, lookahead="0_0_0", chunk_size_left="0", chunk_size_current="0", chunk_size_right="0",
            streaming_type='mask')
CONV = dict(input_dim=80, in_channel=1, channels="32_32", kernel_sizes="(3,3)_(3,3)", strides="(1,1)_(1,1)",
            poolings="(2,2)_(2,2)", dropout=0.0, normalization='', residual=False, bottleneck_dim=32, param_init=0.1)

# name: (args overrides, conv overrides or None, kind, chunk sizes (N_l, N_c, N_r) of the schedule, T)
CASES = {
    # unidirectional Conformer behind the 1/4 CNN: lookback / lookahead trimming, causal attention over the growing
    # cache, causal depthwise conv with its own cache
    "stream_uni_conformer": (dict(enc_type='conv_uni_conformer'), {}, 'conformer', (0, 8, 0), 150),
    # latency-controlled Conformer, chunk-wise mask, hierarchical subsampling (per-layer cache sizes differ)
    "stream_lc_mask_conformer": (dict(chunk_size_left="16", chunk_size_current="8", subsample="2_2_1",
                                      streaming_type='mask'), dict(poolings="(1,1)_(2,2)"), 'conformer', (16, 8, 0), 141),
    # latency-controlled Transformer (relative_xl), overlapped windows
    "stream_lc_reshape_transformer": (dict(enc_type='conv_transformer', pe_type='relative_xl', ffn_activation='relu',
                                           chunk_size_left="8", chunk_size_current="16", chunk_size_right="8",
                                           streaming_type='reshape', clamp_len=-1), {}, 'transformer', (8, 16, 8), 132),
    # unidirectional Transformer with absolute positions (offset carried across chunks), no CNN
    "stream_uni_transformer_add": (dict(enc_type='uni_transformer', pe_type='add', ffn_activation='relu'), None,
                                   'transformer', (0, 4, 0), 61),
    # C4 in miniature: CNN (1/4) + unidirectional LSTM stack with projections, (h, c) carried across chunks of 8 frames
    "stream_conv_lstm": (dict(RNN=True, enc_type='conv_lstm', n_units=32, n_projs=16, n_layers=3, subsample="1_1_1",
                              last_proj_dim=24), {}, 'rnn', (0, 8, 0), 150),
    # latency-controlled BLSTM (N_c = 32, N_r = 16) behind a 1/2 CNN, hierarchical subsampling, summed directions
    "stream_lc_blstm": (dict(RNN=True, enc_type='conv_blstm', n_units=24, n_layers=3, subsample="1_2_1",
                             chunk_size_current="32", chunk_size_right="16", bidir_sum_fwd_bwd=True),
                        dict(channels="32", kernel_sizes="(3,3)", strides="(1,1)", poolings="(2,2)"), 'rnn', (0, 32, 16), 175),
}

RNN_BASE = dict(input_dim=80, enc_type='blstm', n_units=16, n_projs=0, last_proj_dim=0, n_layers=2, n_layers_sub1=0,
                n_layers_sub2=0, dropout_in=0.0, dropout=0.0, subsample="1_1", subsample_type='drop', n_stacks=1,
                n_splices=1, frontend_conv=None, bidir_sum_fwd_bwd=False, task_specific_layer=False, param_init=0.1,
                chunk_size_current="0", chunk_size_right="0", cnn_lookahead=True, rsp_prob=0)


def build(name):
    import importlib
    ov, conv_ov, kind, _, _ = CASES[name]
    ov = dict(ov)
    args = dict(RNN_BASE) if ov.pop("RNN", False) else dict(BASE)
    args.update(ov)
    torch.manual_seed(0)
    conv_args = None
    if conv_ov is not None:
        conv_args = dict(CONV)
        conv_args.update(conv_ov)
        conv_args["bottleneck_dim"] = args["d_model"] if kind != 'rnn' else 0
        conv_mod = importlib.import_module('neural_sp.models.seq2seq.encoders.conv')
        args["frontend_conv"] = conv_mod.ConvEncoder(**conv_args)
    if kind == 'rnn':
        enc = importlib.import_module('neural_sp.models.seq2seq.encoders.rnn').RNNEncoder(**args)
    elif kind == 'conformer':
        enc = importlib.import_module('neural_sp.models.seq2seq.encoders.conformer').ConformerEncoder(**args)
    else:
        a = dict(args)
        a.pop("kernel_size"), a.pop("normalization")
        enc = importlib.import_module('neural_sp.models.seq2seq.encoders.transformer').TransformerEncoder(**a)
    return enc.eval(), args, conv_args, kind


def main():
    for name, (_, _, _, (N_l, N_c, N_r), T) in CASES.items():
        enc, args, conv_args, kind = build(name)
        unidir = 'uni' in args['enc_type'] or args['enc_type'] in ('lstm', 'conv_lstm')
        st = getattr(enc, 'streaming_type', '')
        if st == 'mask':
            N_l = 0
        factor = enc.subsampling_factor
        if kind == 'rnn':
            conv_context = enc.conv.context_size if enc.conv is not None else 0
        else:
            conv_context = enc.conv.context_size if (enc.conv is not None and not enc.lc_bidir) else 0
        rng = np.random.default_rng(4321)
        xs = rng.standard_normal((1, T, 80)).astype(np.float32)
        if st == 'mask' and enc.conv is not None and T % N_c != 0:
            xs = np.concatenate([xs, np.zeros((1, N_c - T % N_c, 80), np.float32)], axis=1)
        xmax = xs.shape[1]
        xs_t = torch.from_numpy(xs)
        enc.reset_cache()
        with torch.no_grad():
            out_all = enc(xs_t.clone(), torch.IntTensor([xmax]), task='all')['ys']
        save = {"sd." + k: v.numpy() for k, v in enc.state_dict().items()}
        sched, ck_lens = [], []
        j = 0
        enc.reset_cache()
        for ci in range(math.ceil(xmax / N_c)):
            start, end = j - N_l - conv_context, (j + N_c + N_r) + conv_context
            chunk = xs_t[:, max(0, start):end]
            pl = pr = 0
            if st == 'reshape':
                xlen = max(factor, min(xmax - j, N_c))
                pl = max(0, -start)
                pr = max(0, end - xmax) if end >= xmax else 0
                chunk = torch.cat([chunk.new_zeros(1, pl, 80), chunk, chunk.new_zeros(1, pr, 80)], dim=1)
            else:
                xlen = max(factor, chunk.size(1))
            lookback = start >= 0 and conv_context > 0
            lookahead = end < xmax and conv_context > 0
            with torch.no_grad():
                o = enc(chunk.clone(), torch.IntTensor([xlen]), task='all', streaming=True, lookback=lookback,
                        lookahead=lookahead)['ys']
            save["ck.%d" % ci] = o['xs'].numpy()
            ck_lens.append(int(o['xlens'][0]))
            sched.append([max(0, start), min(end, xmax), pl, pr, xlen, int(lookback), int(lookahead)])
            j += N_c
            if j > xmax or (not lookahead and conv_context > 0 and unidir):
                break
        cfg = {k: v for k, v in args.items() if k != "frontend_conv"}
        save.update(xs=xs, ys=out_all['xs'].numpy(), ys_lens=out_all['xlens'].numpy().astype(np.int32),
                    sched=np.array(sched, np.int32), ck_lens=np.array(ck_lens, np.int32),
                    cfg=np.array(json.dumps(dict(args=cfg, conv=conv_args, kind=kind))))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
        print(name, "offline", tuple(out_all['xs'].shape), "chunks", len(sched), "first", save["ck.0"].shape)


if __name__ == "__main__":
    main()
