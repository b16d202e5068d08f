#!/usr/bin/env python3
"""build_encoder(args) contract from the UNMODIFIED reference (run in the build container):
    python tests/golden/gen_factory_contract.py  ->  tests/golden/factory_contract.json
For several argument namespaces: class name, state_dict keys with shapes, and the properties callers read
(output_dim, subsampling_factor, ...), and a digest of the freshly initialised weights under torch.manual_seed(0)
(same init distributions drawn in the same order = bit-identical initial weights).  CPU only (modules are only constructed)."""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

BASE = dict(input_dim=80, input_type='speech', emb_dim=0, enc_type='conv_conformer', dec_type='lstm', n_stacks=1, n_splices=1,
            conv_in_channel=1, conv_channels="32_32", conv_kernel_sizes="(3,3)_(3,3)", conv_strides="(1,1)_(1,1)",
            conv_poolings="(1,1)_(2,2)", conv_normalization='', conv_bottleneck_dim=0, param_init=0.1,
            transformer_enc_n_heads=4, enc_n_layers=3, enc_n_layers_sub1=0, enc_n_layers_sub2=0,
            transformer_enc_d_model=64, transformer_enc_d_ff=128, transformer_ffn_bottleneck_dim=0,
            transformer_enc_pe_type='relative', transformer_layer_norm_eps=1e-12, transformer_dec_d_model=96,
            dropout_in=0.0, dropout_enc=0.0, dropout_att=0.0, dropout_enc_layer=0.0, subsample="1_2_1",
            subsample_type='max_pool', task_specific_layer=False, transformer_param_init='xavier_uniform',
            transformer_enc_clamp_len=10, transformer_enc_lookaheads="0_0_0", lc_chunk_size_left="0",
            lc_chunk_size_current="0", lc_chunk_size_right="0", lc_type='reshape', conformer_kernel_size=7,
            conformer_normalization='layer_norm', transformer_ffn_activation='relu', enc_n_units=32, enc_n_projs=0,
            bidirectional_sum_fwd_bwd=False, cnn_lookahead=True, rsp_prob_enc=0.0)

CASES = {
    "conv_conformer": {},
    "conv_conformer_xl_sub": dict(transformer_enc_pe_type='relative_xl', enc_n_layers_sub1=2, dec_type='transformer'),
    "conv_transformer": dict(enc_type='conv_transformer', transformer_enc_pe_type='relative_xl', conv_poolings="(2,2)_(2,2)"),
    "transformer_embed": dict(enc_type='transformer', transformer_enc_pe_type='none', subsample="1_1_1"),
    "conv_uni_conformer": dict(enc_type='conv_uni_conformer', transformer_enc_lookaheads="1_0_1"),
    "blstm": dict(enc_type='blstm', subsample="1_2_1", enc_n_projs=16, bidirectional_sum_fwd_bwd=True),
    "conv_lstm": dict(enc_type='conv_lstm', conv_poolings="(2,2)_(2,2)", subsample="1_1_1", dec_type='transformer'),
    "conv_only": dict(enc_type='conv', conv_bottleneck_dim=48),
}
PROPS = ["output_dim", "output_dim_sub1", "output_dim_sub2", "subsampling_factor", "subsampling_factor_sub1",
         "subsampling_factor_sub2", "enc_type"]


def contract(build_encoder):
    res = {}
    for name, ov in CASES.items():
        a = dict(BASE)
        a.update(ov)
        import hashlib
        import torch
        torch.manual_seed(0)
        enc = build_encoder(argparse.Namespace(**a))
        h = hashlib.sha256()
        for k, v in sorted(enc.state_dict().items()):
            h.update(k.encode())
            h.update(v.detach().cpu().contiguous().numpy().tobytes())
        entry = dict(cls=type(enc).__name__, init_digest=h.hexdigest(),
                     state={k: list(v.shape) for k, v in enc.state_dict().items()},
                     props={})
        for p in PROPS:
            try:
                v = getattr(enc, p)
                entry["props"][p] = v if isinstance(v, str) else int(v)      # numpy ints -> plain ints
            except Exception as e:      # noqa: BLE001  (the reference raises for missing sub-task dims)
                entry["props"][p] = "<%s>" % type(e).__name__
        res[name] = entry
    return res


if __name__ == "__main__":
    from oracle.ref_import import import_reference
    import_reference()
    from neural_sp.models.seq2seq.encoders.build import build_encoder
    json.dump(contract(build_encoder), open(os.path.join(HERE, "factory_contract.json"), "w"), indent=1, sort_keys=True)
    print("wrote factory_contract.json")
