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
    config.addinivalue_line("markers", "experimental: opt-in code paths that have not been validated on hardware yet; "
                                       "run with NSP_EXPERIMENTAL=1")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("NSP_EXPERIMENTAL") == "1":
        return
    skip = pytest.mark.skip(reason="opt-in path not validated on hardware yet: set NSP_EXPERIMENTAL=1 to run")
    for item in items:
        if "experimental" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def split_labels(cat, ylens):
    ys, o = [], 0
    for n in ylens:
        ys.append([int(v) for v in cat[o:o + int(n)]])
        o += int(n)
    return ys
