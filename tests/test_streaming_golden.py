"""Chunk-by-chunk streaming against the reference's streamed outputs (tests/golden/stream_*.npz), CPU part: the encoders'
host logic with the ops replaced by their torch restatements (tests/ops_doubles.py).  The fixtures travel, so this runs
wherever the repository does; tests/test_zz_streaming_gpu.py replays the same schedules through the CUDA kernels."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from enc_util import build_ours, replay_stream

STREAM_CASES = ["stream_uni_conformer", "stream_lc_mask_conformer", "stream_lc_reshape_transformer",
                "stream_uni_transformer_add", "stream_conv_lstm", "stream_lc_blstm"]


@pytest.mark.parametrize("name", STREAM_CASES)
def test_streamed_chunks_match_reference_fixture(name, monkeypatch):
    import ops_doubles
    ops_doubles.install(monkeypatch)
    g = load_golden(name + ".npz")
    enc = build_ours(g, torch.device("cpu"), "fp32")
    outs, lens = replay_stream(enc, g, torch.device("cpu"))
    assert lens == g["ck_lens"].tolist()
    for i, o in enumerate(outs):
        ref = torch.from_numpy(g["ck.%d" % i])
        assert o.shape == ref.shape, (i, o.shape, ref.shape)
        assert torch.allclose(o, ref, atol=1e-4), (i, float((o - ref).abs().max()))
    # offline output of the same module
    enc.reset_cache()
    off = enc(torch.from_numpy(g["xs"]), torch.IntTensor([g["xs"].shape[1]]), task='all')['ys']
    assert off['xlens'].tolist() == g["ys_lens"].tolist()
    assert np.allclose(off['xs'].numpy(), g["ys"], atol=1e-4)
