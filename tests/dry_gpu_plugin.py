"""Debugging aid (not loaded by default): run the `-m gpu` tests in a GPU-less container as a DRY run.

    python -m pytest tests -m gpu -p dry_gpu_plugin -q            # add NSP_EXPERIMENTAL=1 for the opt-in tests

Every `device='cuda...'` a test or the package asks for becomes the CPU (a TorchFunctionMode rewrites device arguments),
torch.cuda's stream / event / synchronisation calls become no-ops, and the library object behind `neural_sp_b200.ops` is the
validating dry library of tests/dry_lib.py (real wrappers, real argument validation of the entry points, no kernels).
Outputs are uninitialised memory, so every numeric comparison fails with an AssertionError raised in the TEST file -- those
are expected and meaningless.  Anything else (TypeError, NspError "bad argument", AttributeError, an assertion inside the
package) is a host-side bug that the GPU run would hit too: that is what this mode is for.  The conftest hook below
reports the split at the end of the session.
"""
import os
import sys

import pytest
import torch
from torch.overrides import TorchFunctionMode

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

_CPU = torch.device("cpu")


def _fix(v):
    if isinstance(v, torch.device) and v.type == "cuda":
        return _CPU
    if isinstance(v, str) and v.startswith("cuda"):
        return "cpu"
    return v


class _CudaToCpu(TorchFunctionMode):
    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        name = getattr(func, "__name__", "")
        if name == "cuda":                                  # Tensor.cuda() -> identity
            return args[0]
        if name == "pin_memory":
            return args[0]
        if "device" in kwargs:
            kwargs["device"] = _fix(kwargs["device"])
        if "pin_memory" in kwargs:
            kwargs["pin_memory"] = False
        if name in ("to", "_to_copy", "new_empty", "new_zeros", "new_full", "new_ones", "new_tensor"):
            args = tuple(_fix(a) for a in args)
            if name == "to":
                kwargs.pop("non_blocking", None)
        return func(*args, **kwargs)


class _Stream:
    cuda_stream = 0

    def __getattr__(self, k):
        return lambda *a, **kw: None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _Event:
    def __init__(self, *a, **kw):
        pass

    def record(self, *a, **kw):
        pass

    def synchronize(self):
        pass

    def wait(self, *a, **kw):
        pass

    def query(self):
        return True

    def elapsed_time(self, other):
        return 1.0


_STATE = {}


def pytest_configure(config):
    import dry_lib
    from neural_sp_b200 import ops

    class _MP:
        def setattr(self, o, n, v):
            setattr(o, n, v)
    _STATE["dry"] = dry_lib.install(_MP(), validate=True)
    mode = _CudaToCpu()
    mode.__enter__()
    _STATE["mode"] = mode
    tc = torch.cuda
    tc.is_available = lambda: True
    tc.synchronize = lambda *a, **k: None
    tc.current_stream = lambda *a, **k: _Stream()
    tc.default_stream = lambda *a, **k: _Stream()
    tc.Stream = lambda *a, **k: _Stream()
    tc.stream = lambda s: _Stream()
    tc.Event = _Event
    tc.set_device = lambda *a, **k: None
    tc.current_device = lambda: 0
    tc.device_count = lambda: 1
    tc.empty_cache = lambda: None
    tc.manual_seed_all = lambda *a, **k: None
    tc.get_device_capability = lambda *a, **k: (10, 0)
    tc.get_device_name = lambda *a, **k: "dry-run"
    nn_to = torch.nn.Module.to
    torch.nn.Module.to = lambda self, *a, **k: nn_to(self, *[_fix(x) for x in a], **{kk: _fix(v) for kk, v in k.items()})
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _STATE["outcomes"] = {"expected numeric mismatch": [], "HOST-SIDE ERROR": [], "passed": []}


@pytest.hookimpl(hookwrapper=True)
def pytest_runtest_makereport(item, call):
    outcome = yield
    rep = outcome.get_result()
    if rep.when != "call":
        return
    if rep.passed:
        _STATE["outcomes"]["passed"].append(item.nodeid)
        return
    if rep.failed and call.excinfo is not None:
        tb = call.excinfo.traceback[-1]
        in_test = os.path.basename(str(tb.path)).startswith("test_") or "numpy/testing" in str(tb.path)
        numeric = call.excinfo.errisinstance(AssertionError) and in_test
        numeric = numeric or "reject_cpu_tensors" in item.nodeid     # the device check is what this mode switches off
        key = "expected numeric mismatch" if numeric else "HOST-SIDE ERROR"
        _STATE["outcomes"][key].append("%s :: %s: %s" % (item.nodeid, call.excinfo.typename, str(call.excinfo.value)[:160]))


def pytest_terminal_summary(terminalreporter):
    o = _STATE["outcomes"]
    terminalreporter.write_line("")
    terminalreporter.write_line("dry GPU run: %d reached their numeric comparison (values are garbage: failures expected), "
                                "%d passed, %d HOST-SIDE ERRORS" % (len(o["expected numeric mismatch"]), len(o["passed"]),
                                                                   len(o["HOST-SIDE ERROR"])))
    for line in o["HOST-SIDE ERROR"]:
        terminalreporter.write_line("  HOST-SIDE: " + line)
    terminalreporter.write_line("entry points called: %d distinct, %d calls" % (len(_STATE["dry"].calls),
                                                                               sum(_STATE["dry"].calls.values())))
    from neural_sp_b200 import _lib
    import dry_lib
    never = [n for n in _lib.SIGNATURES if n not in _STATE["dry"].calls and not n.endswith("_workspace_bytes")
             and not n.endswith("_launches") and n not in dry_lib.HOST_ONLY]
    terminalreporter.write_line("kernel entry points no selected test reached: %s" % (", ".join(sorted(never)) or "none"))
