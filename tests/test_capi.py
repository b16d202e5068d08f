"""The C-ABI library loads (no GPU needed) and exports every symbol include/nsp_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "nsp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nsp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from neural_sp_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 7
    for n in names:
        assert hasattr(_lib.lib, n), "libnsp_b200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "%s has no ctypes signature in _lib.py" % n
    assert set(_lib.SIGNATURES) == set(names)
    assert _lib.lib.nsp_version() >= 100


def test_workspace_queries_are_pure_host_functions():
    from neural_sp_b200._lib import lib
    assert lib.nsp_ctc_loss_workspace_bytes(32, 125, 60) > 3 * 32 * 125 * 121 * 4
    assert lib.nsp_ctc_loss_workspace_bytes(0, 125, 60) == 0
    assert lib.nsp_ctc_align_workspace_bytes(2, 200, 30) > 0


def test_product_never_imports_oracle():
    """The oracle is test infrastructure; the product path must not route through it."""
    pkg = os.path.join(ROOT, "neural_sp_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(d, f)).read()
                assert "oracle" not in re.sub(r"#.*", "", txt).replace("no CPU or PyTorch fallback", ""), (d, f)
                assert "ops_doubles" not in txt, (d, f)        # the torch restatements of the ops are test doubles only
