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
