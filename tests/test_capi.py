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


def _header_prototypes():
    """{name: (restype, [argtypes])} parsed from include/nsp_b200.h (C types mapped to their ctypes equivalents)."""
    src = open(os.path.join(ROOT, "include", "nsp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(nsp_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", src):
        res, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()

        def ctype(decl, is_return=False):
            d = decl.strip()
            if "*" in d:
                base = d.replace("const", "").split("*")[0].strip()
                if base == "char":
                    return ctypes.c_char_p
                return ctypes.c_void_p
            toks = d.replace("const", "").split()
            base = " ".join(toks) if is_return or len(toks) == 1 else " ".join(toks[:-1])
            return {"int": ctypes.c_int, "nsp_status": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float,
                    "size_t": ctypes.c_size_t, "unsigned long long": ctypes.c_uint64, "long long": ctypes.c_longlong, "uint64_t": ctypes.c_uint64,
                    "uint32_t": ctypes.c_uint32,
                    "double": ctypes.c_double, "int32_t": ctypes.c_int32, "unsigned": ctypes.c_uint,
                    "unsigned int": ctypes.c_uint}[base]

        argt = [] if args in ("", "void") else [ctype(a) for a in args.split(",")]
        out[name] = (ctype(res, True), argt)
    return out


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype in include/nsp_b200.h, argument by argument, against the ctypes table the Python side calls
    through (a missing or mistyped argument would not fail at load time: it would corrupt the call on the GPU)."""
    from neural_sp_b200 import _lib
    protos = _header_prototypes()
    assert set(protos) == set(_lib.SIGNATURES)
    is_ptr = lambda t: t in (ctypes.c_void_p, ctypes.c_char_p) or issubclass(t, ctypes._Pointer)

    def same(a, b):
        if is_ptr(a) or is_ptr(b):
            return is_ptr(a) and is_ptr(b)
        return a is b or (ctypes.sizeof(a) == ctypes.sizeof(b) and a._type_ == b._type_)   # c_int vs c_int32
    for name, (res, args) in protos.items():
        sres, sargs = _lib.SIGNATURES[name]
        assert same(res, sres), (name, res, sres)
        assert len(args) == len(sargs), "%s: header has %d arguments, ctypes table %d" % (name, len(args), len(sargs))
        for i, (a, b) in enumerate(zip(args, sargs)):
            assert same(a, b), "%s argument %d: header %s, ctypes %s" % (name, i, a, b)
