"""TEST INFRASTRUCTURE ONLY -- locate and import the unmodified reference (hirofumi0810/neural_sp).

The reference is pure Python/PyTorch, so it runs on CPU in the build container from
/root/reference.  It cannot travel to the GPU box, so this module is used only by
tests/golden/gen_golden.py (fixture generation) and by `-m "not gpu"` tests that
re-validate the oracle restatement when the reference tree is present.
Nothing in neural_sp_b200/ may import this file.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("NSP_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_stubs")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "neural_sp"))


def import_reference():
    """Put the reference and the matplotlib/omegaconf stubs on sys.path; return the package."""
    if not reference_available():
        raise ImportError("reference tree not found at %s" % REFERENCE_ROOT)
    for p in (_STUBS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import neural_sp  # noqa: F401
    return neural_sp
