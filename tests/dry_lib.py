"""Dry run of the REAL `neural_sp_b200/ops.py` wrappers on CPU tensors (test infrastructure only).

`install(monkeypatch)` replaces the library object the wrappers call through by ctypes CALLBACKS built from the very
prototype table the product uses (`_lib.SIGNATURES`, itself pinned to include/nsp_b200.h by tests/test_capi.py): every
kernel entry point becomes a function that converts its arguments exactly as the real call would, records them, and
returns NSP_OK without touching memory; the pure host functions (workspace sizes, version) stay the real ones.  Output
tensors therefore hold whatever `torch.empty` gave -- no values are checked.  What IS checked, for every call the host
code makes on its way through a forward / backward pass:
  * the wrapper's own argument validation (dtypes, ranks, strides, contiguity, shape agreement) passes for the tensors
    the encoder / autograd code hands it -- the ops' torch restatements (ops_doubles.py) do not run these checks;
  * the ctypes conversion succeeds (argument count, ints where the ABI takes ints, pointers where it takes pointers);
  * tensor arguments are alive and large enough (optional per-call hook `inspect`).
`install(monkeypatch, validate=True)` additionally forwards every call to the REAL entry point.  Without a GPU the first
CUDA runtime call inside it fails, so the function returns NSP_ERR_CUDA / NSP_ERR_NO_DEVICE -- but only AFTER its own
host-side argument validation (NSP_CHECK_ARG, the unsupported-shape checks, workspace-size checks) has accepted the call:
a return of NSP_ERR_INVALID or NSP_ERR_UNSUPPORTED is therefore a genuine rejection of what the host code passed and is
raised as an error.  (No entry point dereferences its pointer arguments on the host.)
This is the closest a GPU-less container gets to running the product path: the only code not executed is the kernels.
"""
import collections
import ctypes

HOST_ONLY = ("nsp_version", "nsp_last_error", "nsp_set_gemm_epilogue", "nsp_get_gemm_epilogue")


class DryLib:
    def __init__(self, real, signatures, validate=False):
        self.calls = collections.Counter()
        self.real, self.validate = real, validate
        self.rejected = []
        self._keep = []
        for name, (res, args) in signatures.items():
            fn = getattr(real, name)
            if name.endswith("_workspace_bytes") or name.endswith("_launches") or name in HOST_ONLY:
                setattr(self, name, fn)
                continue
            proto = ctypes.CFUNCTYPE(res, *args)
            cb = proto(self._make(name, len(args)))
            self._keep.append(cb)
            setattr(self, name, cb)

    def _make(self, name, nargs):
        def f(*a):
            assert len(a) == nargs, (name, len(a), nargs)
            self.calls[name] += 1
            if self.validate:
                st = getattr(self.real, name)(*a)
                msg = self.real.nsp_last_error().decode("utf-8", "replace") if st else ""
                # NSP_ERR_INVALID / NSP_ERR_UNSUPPORTED: the argument checks said no (a missing driver entry point for the
                # TMA descriptors is reported as UNSUPPORTED too: that one is the GPU-less container, not the call)
                if st in (1, 3) and "from the driver" not in msg:
                    self.rejected.append((name, st, msg))
                    return st               # -> ops.check raises NspError with the library's message
            return 0
        return f


def install(monkeypatch, validate=False):
    import torch
    from neural_sp_b200 import _lib, ops
    if validate and torch.cuda.is_available() and not getattr(torch.cuda.is_available, "__name__", "") == "<lambda>":
        # with a real device the forwarded calls would LAUNCH kernels on host pointers (and poison the CUDA context)
        raise RuntimeError("dry_lib.install(validate=True) is for GPU-less machines only")
    dry = DryLib(_lib.lib, _lib.SIGNATURES, validate)
    monkeypatch.setattr(ops, "lib", dry)
    monkeypatch.setattr(ops, "_require_cuda", lambda *ts: None)
    monkeypatch.setattr(ops, "current_stream_ptr", lambda: None)
    return dry
