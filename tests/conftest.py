import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def split_labels(cat, ylens):
    ys, o = [], 0
    for n in ylens:
        ys.append([int(v) for v in cat[o:o + int(n)]])
        o += int(n)
    return ys
