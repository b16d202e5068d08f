"""Command-line contract of the encoders (add_args / define_name) against tables extracted from the
unmodified reference (tests/golden/cli_contract.json, generator: tests/golden/gen_cli_contract.py).  CPU only."""
import json
import os
import sys

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)


def test_add_args_and_define_name_match_reference():
    from gen_cli_contract import contract
    from neural_sp_b200.encoders.conformer import ConformerEncoder
    from neural_sp_b200.encoders.conv import ConvEncoder
    from neural_sp_b200.encoders.rnn import RNNEncoder
    from neural_sp_b200.encoders.transformer import TransformerEncoder
    ours = contract({"conformer": ConformerEncoder, "transformer": TransformerEncoder, "rnn": RNNEncoder, "conv": ConvEncoder})
    ref = json.load(open(os.path.join(GOLDEN, "cli_contract.json")))
    assert set(ours) == set(ref)
    for fam in ref:
        assert ours[fam]["options"] == ref[fam]["options"], fam
        assert ours[fam]["names"] == ref[fam]["names"], fam


def test_build_encoder_factory_matches_reference():
    """Same class, same state_dict keys and shapes, same caller-visible properties for every encoder family the factory
    serves (tests/golden/factory_contract.json from the reference's build_encoder)."""
    from gen_factory_contract import contract
    from neural_sp_b200.encoders.build import build_encoder
    ours = contract(build_encoder)
    ref = json.load(open(os.path.join(GOLDEN, "factory_contract.json")))
    assert set(ours) == set(ref)
    for name in ref:
        assert ours[name]["cls"] == ref[name]["cls"], name
        assert ours[name]["state"] == ref[name]["state"], (name, set(ours[name]["state"]) ^ set(ref[name]["state"]))
        assert ours[name]["props"] == ref[name]["props"], (name, ours[name]["props"], ref[name]["props"])
        # same initialisers drawn in the same order: bit-identical fresh weights under the same seed
        assert ours[name]["init_digest"] == ref[name]["init_digest"], name


def test_length_arithmetic_matches_reference():
    """Front-end and subsampler length updates for every input length 1..260 (tests/golden/lens_contract.json: the
    reference's update_lens_1d/2d and subsampler forward passes, including their floor-division quirks at tiny lengths)."""
    import torch
    from neural_sp_b200.encoders import subsampling as S
    from neural_sp_b200.encoders.conv import ConvEncoder
    ref = json.load(open(os.path.join(GOLDEN, "lens_contract.json")))
    lens = list(range(1, 261))
    for pool, want in ref["conv"].items():
        enc = ConvEncoder(80, in_channel=1, channels="32_32", kernel_sizes="(3,3)_(3,3)", strides="(1,1)_(1,1)", poolings=pool,
                          dropout=0., normalization='', residual=False, bottleneck_dim=0, param_init=0.1)
        got = [int(enc.output_lens(torch.IntTensor([n]))[0]) for n in lens]
        assert got == want, pool
    make = {"max_pool": S.MaxPoolSubsampler, "mean_pool": S.MeanPoolSubsampler, "drop": S.DropSubsampler, "add": S.AddSubsampler,
            "concat": lambda f: S.ConcatSubsampler(f, 8), "conv1d": lambda f: S.Conv1dSubsampler(f, 8)}
    for key, want in ref["sub"].items():
        typ, f = key.split("/")
        m = make[typ](int(f))
        got = [int(m._lens(torch.IntTensor([n]))[0]) for n in lens]
        ok = [g == w for g, w in zip(got, want) if w is not None]      # None: the reference itself raises (empty concat)
        assert all(ok), (key, [(n, g, w) for n, g, w in zip(lens, got, want) if w is not None and g != w][:5])
