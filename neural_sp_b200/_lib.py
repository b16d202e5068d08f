"""ctypes binding of libnsp_b200.so (the C ABI declared in include/nsp_b200.h).

The CUDA library is the product: there is no CPU or PyTorch fallback.  Importing this module
without a built library raises, and every call checks the returned nsp_status.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NSP_LIB_PATH selects another build of the SAME library (e.g. libnsp_b200_dbg.so, `make -C neural_sp_b200/csrc debug`:
# bounded mbarrier waits for kernel bring-up); it is not a fallback mechanism -- a missing file still raises.
LIB_PATH = os.environ.get("NSP_LIB_PATH") or os.path.join(_HERE, "libnsp_b200.so")


class NspError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "neural_sp_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C neural_sp_b200/csrc`). There is no CPU fallback." % LIB_PATH)
    return ctypes.CDLL(LIB_PATH)


lib = _load()

c_int, c_i64, c_f32, c_vp, c_sz = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol declared in include/nsp_b200.h
SIGNATURES = {
    "nsp_version": (c_int, []),
    "nsp_last_error": (ctypes.c_char_p, []),
    "nsp_device_info": (c_int, [ctypes.POINTER(c_int)] * 3),
    "nsp_ctc_loss_workspace_bytes": (c_sz, [c_int, c_int, c_int]),
    "nsp_ctc_loss_fwd_bwd": (c_int, [c_vp, c_i64, c_i64, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp,
                                     c_int, c_f32, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "nsp_ctc_align_workspace_bytes": (c_sz, [c_int, c_int, c_int]),
    "nsp_ctc_forced_align": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_int, c_vp,
                                     c_vp, c_sz, c_vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)   # AttributeError here = header/library mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args


def check(status, what=""):
    if status != 0:
        msg = lib.nsp_last_error().decode("utf-8", "replace")
        raise NspError("%s failed with nsp_status=%d: %s" % (what or "nsp call", status, msg))


def ptr(t):
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

# ---- GEMM / elementwise ----
_more = {
    "nsp_linear_fwd": (c_int, [c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_int,
                               c_vp, c_vp, c_i64, c_f32, c_vp, c_i64, c_int, c_vp, c_i64, c_vp]),
    "nsp_split_tf32": (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "nsp_cast_f32_to_bf16": (c_int, [c_vp, c_vp, c_i64, c_vp]),
}
for _name, (_res, _args) in _more.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args
SIGNATURES.update(_more)

_more = {
    "nsp_layernorm_fwd": (c_int, [c_vp, c_i64, c_vp, c_vp, c_f32, c_f32, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_vp]),
    "nsp_relpos_attention_fwd": (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_vp,
                                         c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                         c_int, c_int, c_vp]),
    "nsp_conformer_conv_fwd": (c_int, [c_int, c_vp, c_i64, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_f32,
                                       c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_vp]),
}
for _name, (_res, _args) in _more.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args
SIGNATURES.update(_more)

_more = {
    "nsp_scale_inplace": (c_int, [c_vp, c_f32, c_i64, c_vp]),
    "nsp_mask_rects": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_int, c_vp]),
    "nsp_add_pos_enc": (c_int, [c_vp, c_vp, c_f32, c_int, c_int, c_int, c_vp]),
    "nsp_colsum": (c_int, [c_vp, c_vp, c_int, c_int, c_vp]),
    "nsp_xl_pos_table": (c_int, [c_vp, c_vp, c_int, c_int, c_vp]),
    "nsp_conv3x3_relu_fwd": (c_int, [c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_vp]),
    "nsp_maxpool2d_fwd": (c_int, [c_int, c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "nsp_maxpool_time_fwd": (c_int, [c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
}
for _name, (_res, _args) in _more.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args
SIGNATURES.update(_more)

_more = {
    "nsp_rnnt_loss_workspace_bytes": (c_sz, [c_int, c_int, c_int]),
    "nsp_rnnt_loss_fwd_bwd": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp,
                                      c_vp, c_sz, c_vp]),
}
for _name, (_res, _args) in _more.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args
SIGNATURES.update(_more)

_more = {"nsp_conv3x3_c32_tc_fwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp])}
for _name, (_res, _args) in _more.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args
SIGNATURES.update(_more)

_more = {
    "nsp_softmax_rows": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_f32, c_vp]),
    "nsp_ctc_greedy": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "nsp_rnnt_joint_tanh": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
}
for _name, (_res, _args) in _more.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args
SIGNATURES.update(_more)

_more = {"nsp_pool_time_fwd": (c_int, [c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp])}
for _name, (_res, _args) in _more.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args
SIGNATURES.update(_more)

_more = {"nsp_lstm_workspace_bytes": (c_sz, [c_int, c_int, c_int]),
         "nsp_lstm_tc_supported": (c_int, [c_int, c_int, c_int]),
         "nsp_lstm_tc_workspace_bytes": (c_sz, [c_int, c_int, c_int, c_int]),
         "nsp_lstm_seq_fwd_tc": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                         c_vp, c_vp, c_sz, c_vp]),
         "nsp_lstm_seq_bwd_tc": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp,
                                         c_vp, c_sz, c_vp]),
         "nsp_lstm_seq_fwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_sz, c_vp]),
         "nsp_lstm_seq_fwd_state": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                            c_sz, c_vp]),
         "nsp_lstm_seq_fwd_save_state": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                                 c_vp, c_vp, c_vp, c_sz, c_vp]),
         "nsp_lstm_seq_bwd_state": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp,
                                            c_vp, c_vp, c_sz, c_vp]),
         "nsp_lstm_seq_fwd_save": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
         "nsp_lstm_seq_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_sz, c_vp])}
for _name, (_res, _args) in _more.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args
SIGNATURES.update(_more)

# ---- backward pass ----
_more = {
    "nsp_linear_fwd_save": (c_int, [c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_int,
                                    c_vp, c_vp, c_i64, c_f32, c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "nsp_linear_wgrad": (c_int, [c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_f32, c_vp, c_i64,
                                 c_int, c_vp]),
    "nsp_layernorm_bwd": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_f32, c_f32, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64,
                                  c_vp, c_vp, c_vp, c_f32, c_int, c_int, c_vp]),
    "nsp_act_bwd_bias": (c_int, [c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
    "nsp_act_bwd": (c_int, [c_int, c_int, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "nsp_glu_bwd": (c_int, [c_int, c_vp, c_vp, c_vp, c_i64, c_int, c_vp]),
    "nsp_colsum_acc": (c_int, [c_int, c_vp, c_i64, c_int, c_int, c_f32, c_vp, c_vp]),
    "nsp_maxpool_time_bwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
    "nsp_dropout": (c_int, [c_int, c_int, c_vp, c_vp, c_i64, c_f32, c_f32, c_vp, ctypes.c_uint32, c_vp]),
    "nsp_dropout_add": (c_int, [c_int, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_vp, ctypes.c_uint32, c_vp]),
    "nsp_rng_advance": (c_int, [c_vp, c_vp]),
    "nsp_rnnt_grad_logits": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_sz, c_vp, c_vp,
                                     c_int, c_vp]),
    "nsp_dwconv_stats_fwd": (c_int, [c_int, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "nsp_bn_bwd": (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_i64, c_i64, c_int, c_vp]),
    "nsp_bn_swish_bwd": (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_i64, c_i64, c_int,
                                 c_vp]),
    "nsp_gn2_swish_bwd": (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_f32, c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_vp]),
    "nsp_dwconv_bwd": (c_int, [c_int, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int,
                               c_vp]),
    "nsp_log_softmax_bwd": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_vp]),
    "nsp_rnnt_joint_tanh_bwd": (c_int, [c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
    "nsp_pool_time_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "nsp_relu_mask": (c_int, [c_int, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "nsp_maxpool2d_relu_bwd": (c_int, [c_int, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "nsp_conv3x3_c32_wgrad_tc": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp]),
    "nsp_conv3x3_wgrad": (c_int, [c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "nsp_relpos_attention_bwd_workspace_bytes": (c_sz, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "nsp_relpos_attention_bwd": (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_vp, c_vp,
                                         c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64,
                                         c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                         c_vp, c_sz, c_vp]),
    "nsp_relpos_attention_fwd_stats": (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_vp,
                                               c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                               c_int, c_int, c_vp, ctypes.POINTER(c_int), c_vp]),
    "nsp_conformer_conv_bwd": (c_int, [c_int, c_vp, c_i64, c_vp, c_vp, c_int, c_vp, c_vp, c_f32, c_vp, c_i64, c_vp, c_i64,
                                       c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_sz, c_vp]),
    "nsp_conformer_conv_bwd_workspace_bytes": (c_sz, [c_int, c_int, c_int, c_int]),
}
for _name, (_res, _args) in _more.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args
SIGNATURES.update(_more)

_more = {"nsp_set_gemm_epilogue": (c_int, [c_int]), "nsp_get_gemm_epilogue": (c_int, []),
         "nsp_gemm_tma_epilogue_launches": (ctypes.c_longlong, []), "nsp_gemm_cta_pair_launches": (ctypes.c_longlong, []),
         "nsp_wgrad_tma_epilogue_launches": (ctypes.c_longlong, [])}
for _name, (_res, _args) in _more.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args
SIGNATURES.update(_more)

# GEMM kernel family: the library default is "pair" (TMA-store epilogue, CTA pairs / cta_group::2 on the large problems;
# validated and fastest on B200 in round 2).  NSP_GEMM_EPILOGUE=direct | tma pins the older families (tests, A/B runs).
_GEMM_MODES = {"direct": 0, "tma": 1, "pair": 2}
_mode = os.environ.get("NSP_GEMM_EPILOGUE", "").lower()
if _mode and _mode not in _GEMM_MODES:
    raise NspError("NSP_GEMM_EPILOGUE=%r (expected one of %s)" % (_mode, sorted(_GEMM_MODES)))
if _mode:
    check(lib.nsp_set_gemm_epilogue(_GEMM_MODES[_mode]), "nsp_set_gemm_epilogue")
