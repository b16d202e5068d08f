#!/usr/bin/env python3
"""Fixtures for features added after the round's GPU budget was spent (consumed by tests/test_zz_*.py on the GPU and by the
CPU wiring tests): same content as enc_*.npz / encgrad_*.npz (see gen_golden_encoder.py), under a `zz_` prefix so that the
already-validated GPU test files (which glob enc_*.npz / encgrad_*.npz) keep their case lists.

    python tests/golden/gen_golden_zz.py"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from oracle.ref_import import import_reference  # noqa: E402

import_reference()
import gen_golden_encoder as G  # noqa: E402

ZZ_CASES = {
    # Conformer v2 blocks (FFN -> conv -> plain MHA -> FFN), hierarchical max-pool, LayerDrop rescale in eval
    "conformer_v2": dict(args=dict(enc_type='conv_conformer_v2', dropout_layer=0.0), conv={}, B=3, T=64, xlens=[64, 57, 40]),
    # unidirectional v2: causal attention + causal depthwise conv
    "uni_conformer_v2": dict(args=dict(enc_type='conv_uni_conformer_v2', n_layers=2, subsample="1_1", lookahead="0_0",
                                       kernel_size=5), conv=dict(poolings="(2,2)_(2,2)"), B=2, T=60, xlens=[60, 48]),
    # plain Transformer with absolute positions (pe_type='add'), GELU FFN, bridge
    "transformer_add": dict(args=dict(enc_type='conv_transformer', pe_type='add', ffn_activation='gelu', n_layers=2,
                                      subsample="1_2", lookahead="0_0", last_proj_dim=40),
                            conv=dict(poolings="(2,2)_(2,2)"), B=2, T=70, xlens=[70, 51], kind='transformer'),
    # no CNN: embedding Linear + absolute positions, plain MHA
    "transformer_embed_add": dict(args=dict(enc_type='transformer', pe_type='add', ffn_activation='relu', n_layers=2,
                                            subsample="1_2", lookahead="0_0", n_heads=2, d_model=32, d_ff=64),
                                  conv=None, B=2, T=50, xlens=[50, 44], kind='transformer'),
    # low-rank feed-forward (ffn_bottleneck_dim > 0: w_1_e / w_1_d / w_2_e / w_2_d), Conformer and Transformer + GLU
    "conformer_lowrank": dict(args=dict(ffn_bottleneck_dim=16, n_layers=2, subsample="1_1", lookahead="0_0", d_model=32, d_ff=64,
                                        n_heads=2), conv=dict(poolings="(2,2)_(2,2)"), B=2, T=60, xlens=[60, 47]),
    "transformer_lowrank_glu": dict(args=dict(enc_type='conv_transformer', pe_type='relative_xl', ffn_activation='glu',
                                              ffn_bottleneck_dim=24, n_layers=2, subsample="1_1", lookahead="0_0", d_model=32,
                                              d_ff=64, n_heads=2), conv=dict(poolings="(2,2)_(2,2)"), B=2, T=60, xlens=[60, 47],
                                    kind='transformer'),
    # task-specific extra block on the sub-task output (sub_module, transformer.py:619-631) + bridge + sub-task LayerNorm
    "conformer_task_specific": dict(args=dict(n_layers=3, n_layers_sub1=2, task_specific_layer=True, subsample="1_2_1",
                                              d_model=32, d_ff=64, n_heads=2, last_proj_dim=24),
                                    conv=dict(poolings="(2,2)_(2,2)"), B=2, T=60, xlens=[60, 47]),
    # GLU feed-forward activation (LinearGLUBlock) + 'drop' subsampling
    "conformer_glu_drop": dict(args=dict(ffn_activation='glu', subsample_type='drop', n_layers=2, subsample="2_1",
                                         lookahead="0_0", d_model=32, d_ff=64, n_heads=2),
                               conv=dict(poolings="(2,2)_(2,2)"), B=2, T=61, xlens=[61, 50]),
    # the remaining hierarchical subsamplers in the training path: concat, conv1d, add, mean_pool
    "transformer_concat": dict(args=dict(enc_type='conv_transformer', pe_type='relative_xl', ffn_activation='relu', n_layers=2,
                                         subsample="2_1", lookahead="0_0", subsample_type='concat', d_model=32, d_ff=64,
                                         n_heads=2), conv=dict(poolings="(2,2)_(2,2)"), B=2, T=66, xlens=[66, 45],
                               kind='transformer'),
    "conformer_conv1d": dict(args=dict(subsample_type='conv1d', n_layers=2, subsample="2_1", lookahead="0_0", d_model=32,
                                       d_ff=64, n_heads=2), conv=dict(poolings="(2,2)_(2,2)"), B=2, T=70, xlens=[70, 52]),
    "conformer_add": dict(args=dict(subsample_type='add', n_layers=2, subsample="2_1", lookahead="0_0", d_model=32, d_ff=64,
                                    n_heads=2), conv=dict(poolings="(2,2)_(2,2)"), B=2, T=70, xlens=[70, 52]),
    # (lengths chosen so that the reference's own floor-formula lengths agree with the ceil-mode pooled tensor: it asserts
    #  on its mask shape otherwise, relative_multihead_attention.py:166)
    "conformer_mean": dict(args=dict(subsample_type='mean_pool', n_layers=2, subsample="2_1", lookahead="0_0", d_model=32,
                                     d_ff=64, n_heads=2), conv=dict(poolings="(2,2)_(2,2)"), B=2, T=64, xlens=[64, 56]),
}


def main():
    for short, case in ZZ_CASES.items():
        name = "enc_" + short
        G.CASES[name] = case
        enc, args, conv_args, kind = G.build_reference(name)
        rng = np.random.default_rng(1234)
        B, T = case["B"], case["T"]
        xs = np.zeros((B, T, 80), np.float32)
        for b, n in enumerate(case["xlens"]):
            xs[b, :n] = rng.standard_normal((n, 80)).astype(np.float32)
        out = enc(torch.from_numpy(xs), torch.IntTensor(case["xlens"]), task='all')
        ys = out['ys']['xs']
        save = {"sd." + k: v.numpy() for k, v in enc.state_dict().items()}
        save.update(xs=xs, xlens=np.array(case["xlens"], np.int32), ys=ys.detach().numpy(),
                    xlens_out=out['ys']['xlens'].numpy().astype(np.int32))
        cfg = {k: v for k, v in args.items() if k != "frontend_conv"}
        save["cfg"] = np.array(json.dumps(dict(args=cfg, conv=conv_args, kind=kind)))
        if out['ys_sub1']['xs'] is not None:
            save["ys_sub1"] = out['ys_sub1']['xs'].detach().numpy()
        np.savez_compressed(os.path.join(HERE, "zz_" + name + ".npz"), **save)
        w = torch.from_numpy(G.grad_loss_weights(tuple(ys.shape), out['ys']['xlens'].tolist()))
        loss = (ys * w).sum()
        if out['ys_sub1']['xs'] is not None:      # sub-task output: stored and part of the loss (weights seeded 99)
            s1 = out['ys_sub1']['xs']
            loss = loss + (s1 * torch.from_numpy(G.grad_loss_weights(tuple(s1.shape), out['ys_sub1']['xlens'].tolist(),
                                                                      seed=99))).sum()
        loss.backward()
        gsave = {"g." + k: p.grad.numpy() for k, p in enc.named_parameters() if p.grad is not None}
        gsave["loss"] = np.array(float(loss.detach()), np.float32)
        np.savez_compressed(os.path.join(HERE, "zz_encgrad_" + short + ".npz"), **gsave)
        print(short, tuple(ys.shape), float(loss), len(gsave) - 1, "gradient tensors; no grad:",
              [k for k, p in enc.named_parameters() if p.grad is None])


if __name__ == "__main__":
    main()
