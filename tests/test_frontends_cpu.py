"""Input-side helpers: SpecAugment with the reference's random draws, and the one-copy host->device staging.
CPU part: seeded runs of neural_sp_b200.frontends.spec_augment.SpecAugment pick the SAME rectangles and produce the same masked
batch as the unmodified reference class (all presets, adaptive variant); pad_and_upload equals the reference's pad_list.
Needs /root/reference for the comparisons (skipped elsewhere); the mask kernel itself is checked in test_zz_conv_frontend_gpu.py."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
ref_only = pytest.mark.skipif(not os.path.isdir("/root/reference/neural_sp"), reason="reference tree not available")


@ref_only
@pytest.mark.parametrize("cfg", [dict(F=27, T=100, n_freq_masks=2, n_time_masks=2), dict(F=15, T=70, n_freq_masks=2, n_time_masks=2, p=0.2),
                                 dict(F=27, T=0, n_freq_masks=1, n_time_masks=0, adaptive_number_ratio=0.04, adaptive_size_ratio=0.04),
                                 dict(F=27, T=100, n_freq_masks=0, n_time_masks=1), "librispeech_double", "switchboard_strong"])
def test_spec_augment_matches_reference_draws(cfg, monkeypatch):
    import ops_doubles
    from oracle.ref_import import import_reference
    import_reference()
    from neural_sp.models.seq2seq.frontends.spec_augment import SpecAugment as RefSA
    from neural_sp_b200.frontends.spec_augment import SpecAugment
    ops_doubles.install(monkeypatch)
    if isinstance(cfg, str):
        ref, ours = RefSA(27, 100, 1, 1), SpecAugment(27, 100, 1, 1)
        getattr(ref, cfg)(), getattr(ours, cfg)()
    else:
        ref, ours = RefSA(**cfg), SpecAugment(**cfg)
    rng = np.random.RandomState(0)
    for trial in range(4):
        xs = torch.from_numpy(rng.randn(3, 300 + 37 * trial, 80).astype(np.float32))
        np.random.seed(100 + trial)
        r = ref(xs.clone())
        np.random.seed(100 + trial)
        o = ours(xs.clone())
        assert torch.equal(r, o)
        assert ours.freq_mask == ref.freq_mask and ours.time_mask == ref.time_mask
        # both consumed the same number of draws from numpy's global generator:
        np.random.seed(100 + trial)
        ref(xs.clone())
        a = np.random.uniform()
        np.random.seed(100 + trial)
        ours(xs.clone())
        assert np.random.uniform() == a


@ref_only
def test_pad_and_upload_matches_pad_list():
    from oracle.ref_import import import_reference
    import_reference()
    from neural_sp.models.torch_utils import np2tensor, pad_list
    from neural_sp_b200.frontends.input import pad_and_upload
    rng = np.random.RandomState(0)
    for _ in range(3):                      # repeated calls reuse (and once grow) the staging buffer
        xs = [rng.randn(n, 80).astype(np.float32) for n in rng.randint(5, 700, size=4)]
        ref = pad_list([np2tensor(x, "cpu").float() for x in xs], 0.)
        dev, lens = pad_and_upload(xs, "cpu")
        assert torch.equal(ref, dev) and lens.tolist() == [len(x) for x in xs] and isinstance(lens, torch.IntTensor)
