"""Pins the oracle (oracle/*.py) to golden vectors produced by the unmodified reference
(tests/golden/gen_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, split_labels
from oracle import ctc_oracle

CTC_CASES = sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLDEN, "ctc_*.npz")) if "head" not in f)


@pytest.mark.parametrize("name", CTC_CASES)
def test_ctc_oracle_matches_reference(name):
    g = load_golden(name)
    ys = split_labels(g["ys_cat"], g["ylens"])
    loss, grad, nll = ctc_oracle.ctc_forward(g["logits"], ys, g["elens"], float(g["lsm"]))
    assert abs(loss - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    np.testing.assert_allclose(nll, g["nll"], rtol=1e-5, atol=1e-3)
    # the reference's gradient is fp32 (ATen); the oracle is fp64
    np.testing.assert_allclose(grad, g["grad"], rtol=0, atol=2e-4)
    if "trigger_points" in g.files:
        trig = ctc_oracle.forced_align(g["logits"], g["elens"], ys)
        assert np.array_equal(trig, g["trigger_points"])     # bit-exact


def test_ctc_oracle_zero_infinity():
    g = load_golden("ctc_infeasible.npz")
    ys = split_labels(g["ys_cat"], g["ylens"])
    nll, loss, grad = ctc_oracle.ctc_nll_and_grad(g["logits"], ys, g["elens"])
    assert nll[1] == 0.0 and nll[2] == 0.0 and nll[0] > 0       # T < L (+repeats) -> zeroed
    assert np.all(grad[1] == 0) and np.all(grad[2] == 0)
