#!/usr/bin/env python3
"""Extract the encoders' command-line contract from the UNMODIFIED reference (run in the build container):
    python tests/golden/gen_cli_contract.py   ->  tests/golden/cli_contract.json
For every encoder family: the argparse options its ``add_args`` registers (flag, type, default, choices) and the
directory names its ``define_name`` produces for a set of option overrides."""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = {
    "conformer": dict(enc_type="conv_conformer", overrides=[
        {}, {"conv_channels": "32_32", "transformer_enc_clamp_len": 10, "conformer_kernel_size": 15,
             "conformer_normalization": "layer_norm", "enc_n_layers": 12},
        {"dropout_enc_layer": 0.1, "lc_chunk_size_left": "64", "lc_chunk_size_current": "64", "lc_type": "mask",
         "transformer_ffn_bottleneck_dim": 128},
        {"enc_type": "conv_uni_conformer", "transformer_enc_lookaheads": "1_0_1_0", "conv_channels": "32_32",
         "conv_normalization": "layer_norm"}]),
    "transformer": dict(enc_type="conv_transformer", overrides=[
        {}, {"conv_channels": "32_32_32", "transformer_enc_pe_type": "relative_xl", "transformer_enc_clamp_len": 5,
             "enc_n_layers": 24, "transformer_enc_d_model": 512},
        {"enc_type": "transformer", "dropout_enc_layer": 0.2, "lc_chunk_size_left": "0_32", "lc_chunk_size_current": "16_32",
         "lc_chunk_size_right": "0_16"}]),
    "rnn": dict(enc_type="conv_blstm", overrides=[
        {}, {"conv_channels": "32_32", "enc_n_units": 1024, "enc_n_projs": 256, "enc_n_layers": 6,
             "bidirectional_sum_fwd_bwd": True},
        {"enc_type": "lstm", "lc_chunk_size_left": "40", "lc_chunk_size_right": "20", "cnn_lookahead": False,
         "rsp_prob_enc": 0.5}]),
    "conv": dict(enc_type="conv", overrides=[{"conv_channels": "32_32", "conv_normalization": "batch_norm"}, {}]),
}


def describe(parser):
    out = []
    for a in parser._actions:
        if not a.option_strings or a.option_strings[0] in ("-h", "--help"):
            continue
        out.append(dict(flag=a.option_strings[0], type=getattr(a.type, "__name__", str(a.type)), default=a.default,
                        choices=list(a.choices) if a.choices is not None else None))
    return sorted(out, key=lambda d: d["flag"])


def contract(classes):
    res = {}
    for fam, case in CASES.items():
        cls = classes[fam]
        ns = argparse.Namespace(enc_type=case["enc_type"])
        parser = argparse.ArgumentParser()
        cls.add_args(parser, ns)
        entry = dict(options=describe(parser), names=[])
        for ov in case["overrides"]:
            args = parser.parse_args([])
            args.enc_type = case["enc_type"]
            args.enc_n_layers = 5
            for k, v in ov.items():
                setattr(args, k, v)
            try:
                name = cls.define_name("", args)
            except AssertionError:
                name = "<AssertionError>"
            entry["names"].append(dict(overrides=ov, name=name))
        res[fam] = entry
    return res


if __name__ == "__main__":
    from oracle.ref_import import import_reference
    import_reference()
    import importlib
    enc = 'neural_sp.models.seq2seq.encoders.'
    classes = {"conformer": importlib.import_module(enc + 'conformer').ConformerEncoder,
               "transformer": importlib.import_module(enc + 'transformer').TransformerEncoder,
               "rnn": importlib.import_module(enc + 'rnn').RNNEncoder,
               "conv": importlib.import_module(enc + 'conv').ConvEncoder}
    json.dump(contract(classes), open(os.path.join(HERE, "cli_contract.json"), "w"), indent=1, sort_keys=True)
    print("wrote cli_contract.json")
