"""The whole `-m gpu` suite (validated + opt-in tests) as a dry run on the CPU (tests/dry_gpu_plugin.py): every test must get
as far as its numeric comparison -- through the real wrappers and the entry points' own argument validation -- without a
host-side error.  Guards the round-end GPU run against Python-side regressions made while no GPU was at hand."""
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.skipif(torch.cuda.is_available(), reason="dry runs are for GPU-less machines: with a device the "
                                "forwarded calls would launch kernels on host pointers")


def test_every_gpu_test_reaches_its_numeric_comparison():
    env = dict(os.environ, NSP_EXPERIMENTAL="1", PYTHONPATH=os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "-p", "dry_gpu_plugin", "-q",
                        "-p", "no:cacheprovider", "--timeout=300"], cwd=os.path.join(ROOT, "tests"), env=env,
                       capture_output=True, text=True, timeout=1500)
    m = re.search(r"dry GPU run: (\d+) reached their numeric comparison.*?(\d+) passed, (\d+) HOST-SIDE ERRORS", r.stdout)
    assert m, r.stdout[-3000:] + r.stderr[-2000:]
    host = [l for l in r.stdout.splitlines() if l.startswith("  HOST-SIDE")]
    assert int(m.group(3)) == 0, "\n".join(host)
    assert int(m.group(1)) >= 400                       # the suite really ran (417 GPU tests at the time of writing)
