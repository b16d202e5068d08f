"""Thin tensor-level wrappers over the C ABI (include/nsp_b200.h).

Each function takes CUDA tensors, allocates outputs/workspaces through torch's caching allocator on
the current stream, and enqueues the CUDA kernels through ctypes.  No arithmetic happens in Python.
"""
import os

import torch

from . import _lib
from ._lib import lib, check, ptr, current_stream_ptr


# ---- launch accounting / optional per-op CUDA-event profiling (used by bench.py) ----
LAUNCHES = 0          # number of CUDA kernels this library launched (claimed count; see KERNELS_PER_CALL)
PROFILE = None        # None, or dict name -> [list of (start_event, end_event)], flops, bytes
KERNELS_PER_CALL = {"nsp_ctc_loss_fwd_bwd": 2, "nsp_ctc_forced_align": 2}


def profile_start():
    global PROFILE
    PROFILE = {}


def profile_stop():
    """-> {name: dict(ms=total, calls=n, flops=sum, bytes=sum)}; synchronises."""
    global PROFILE
    prof, PROFILE = PROFILE, None
    torch.cuda.synchronize()
    out = {}
    for name, rec in (prof or {}).items():
        ms = sum(a.elapsed_time(b) for a, b in rec["ev"])
        out[name] = dict(ms=ms, calls=len(rec["ev"]), flops=rec["flops"], bytes=rec["bytes"])
    return out


SHAPE_TAGS = False    # bench.py --shape-profile: key the GEMM records by shape
LAST_CTC_WS = None


def _run(name, fn, *args, flops=0, nbytes=0, tag=None, shape=None):
    global LAUNCHES
    LAUNCHES += KERNELS_PER_CALL.get(name, 1)
    if PROFILE is None:
        return check(fn(*args), name)
    key = tag or name
    if SHAPE_TAGS and shape is not None:
        key = "%s:%s" % (key, "x".join(str(v) for v in shape))
    rec = PROFILE.setdefault(key, {"ev": [], "flops": 0, "bytes": 0})
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    check(fn(*args), name)
    b.record()
    rec["ev"].append((a, b))
    rec["flops"] += flops
    rec["bytes"] += nbytes


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.NspError("neural_sp_b200 ops need CUDA tensors (no CPU fallback); got device %s" % t.device)


_LABEL_CACHE = {}


def pack_labels(ys, device):
    """List of label lists -> (labels int32 [B, Lmax] on device, ylens int32 [B] on device, Lmax).
    The last packed batch is cached (same labels again -> no new H2D copy; needed for CUDA-graph replays)."""
    key = (tuple(tuple(int(v) for v in y) for y in ys), str(device))
    hit = _LABEL_CACHE.get("last")
    if hit is not None and hit[0] == key:
        return hit[1]
    out = _pack_labels(ys, device)
    _LABEL_CACHE["last"] = (key, out)
    return out


def _pack_labels(ys, device):
    B = len(ys)
    ylens = [len(y) for y in ys]
    Lmax = max(1, max(ylens) if ylens else 1)
    host = torch.zeros(B, Lmax + 1, dtype=torch.int32)
    for b, y in enumerate(ys):
        if len(y):
            host[b, :len(y)] = torch.as_tensor(list(y), dtype=torch.int32)
        host[b, Lmax] = len(y)
    if host.device != device:
        host = host.pin_memory() if torch.cuda.is_available() else host
    dev = host.to(device, non_blocking=True)
    return dev[:, :Lmax].contiguous(), dev[:, Lmax].contiguous(), Lmax


def ctc_loss_fwd_bwd(logits, labels, elens, ylens, blank=0, lsm_prob=0.0):
    """Fused CTC forward+backward (nsp_ctc_loss_fwd_bwd).

    Args:
        logits: fp32 CUDA tensor viewed as [B, T, V] (any b/t strides, unit v stride)
        labels: int32 [B, Lmax] CUDA; elens, ylens: int32 [B] CUDA
    Returns:
        loss (0-dim), nll [B], grad [B, T, V] (contiguous)
    """
    _require_cuda(logits, labels, elens, ylens)
    assert logits.dtype == torch.float32 and logits.dim() == 3 and logits.stride(2) == 1
    B, T, V = logits.shape
    Lmax = labels.shape[1]
    ws_bytes = lib.nsp_ctc_loss_workspace_bytes(B, T, Lmax)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=logits.device)
    nll = torch.empty(B, dtype=torch.float32, device=logits.device)
    loss = torch.empty((), dtype=torch.float32, device=logits.device)
    grad = torch.empty(B, T, V, dtype=torch.float32, device=logits.device)
    _run("nsp_ctc_loss_fwd_bwd", lib.nsp_ctc_loss_fwd_bwd, ptr(logits), logits.stride(0), logits.stride(1), B, T, V,
                                   ptr(labels), Lmax, ptr(elens), ptr(ylens), int(blank), float(lsm_prob),
                                   ptr(nll), ptr(loss), ptr(grad), ptr(ws), ws_bytes, current_stream_ptr(),
         nbytes=8.0 * B * T * V, tag="ctc_loss")
    if os.environ.get("NSP_CTC_DEBUG"):
        global LAST_CTC_WS
        LAST_CTC_WS = ws                  # bring-up: profiles/prof_ctc.py reads the kernel's trace words from the tail
    return loss, nll, grad


def ctc_forced_align(logits, labels, elens, ylens, blank=0):
    """CTC forced alignment (nsp_ctc_forced_align) -> int32 [B, Lmax+1] trigger points."""
    _require_cuda(logits, labels, elens, ylens)
    logits = logits.contiguous()
    assert logits.dtype == torch.float32 and logits.dim() == 3
    B, T, V = logits.shape
    Lmax = labels.shape[1]
    ws_bytes = lib.nsp_ctc_align_workspace_bytes(B, T, Lmax)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=logits.device)
    trig = torch.empty(B, Lmax + 1, dtype=torch.int32, device=logits.device)
    _run("nsp_ctc_forced_align", lib.nsp_ctc_forced_align, ptr(logits), B, T, V, ptr(labels), Lmax, ptr(elens), ptr(ylens), int(blank),
                                   ptr(trig), ptr(ws), ws_bytes, current_stream_ptr())
    return trig


# ---------------------------------------------------------------------------------------------
# tensor-core projections
# ---------------------------------------------------------------------------------------------
PREC = {"bf16": 0, "tf32": 1, "fp32": 2}
ACT = {None: 0, "none": 0, "relu": 1, "swish": 2, "gelu": 3, "gelu_accurate": 4}


def split_tf32(x):
    """fp32 tensor -> (hi, lo) with hi = tf32-rounded x and lo = x - hi (nsp_split_tf32)."""
    _require_cuda(x)
    x = x.contiguous()
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    _run("nsp_split_tf32", lib.nsp_split_tf32, ptr(x), ptr(hi), ptr(lo), x.numel(), current_stream_ptr())
    return hi, lo


def to_bf16(x):
    """fp32 -> bf16 on the library's cast kernel (bf16 input is returned as is)."""
    if x.dtype == torch.bfloat16:
        return x
    _require_cuda(x)
    x = x.contiguous()
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _run("nsp_cast_f32_to_bf16", lib.nsp_cast_f32_to_bf16, ptr(x), ptr(y), x.numel(), current_stream_ptr())
    return y


def _pad_k(t, mult=8):
    """Zero-pad the last (K) dim to a multiple of `mult` elements: TMA needs 16-byte row pitches."""
    K = t.shape[-1]
    Kp = -(-K // mult) * mult
    if Kp == K:
        return t
    return torch.nn.functional.pad(t, (0, Kp - K))


def prepare_weight(w, prec):
    """Operand form of a [N, K] weight for the given precision: bf16 copy, fp32, or (hi, lo) split.
    K is zero-padded to a multiple of 8 so that every row pitch is a 16-byte multiple."""
    w = _pad_k(w.detach())
    if prec == "bf16":
        return (to_bf16(w.detach()),)
    if prec == "tf32":
        return (w.detach().float().contiguous(),)
    return split_tf32(w.detach().float())


def linear(x, w_prepared, bias=None, prec="bf16", act=None, glu=False, residual=None, alpha=1.0,
           out_dtype=torch.float32, out=None, out2_bf16=False, save_pre=False):
    """out = residual + alpha * act(x @ w^T + bias) on the tcgen05 GEMM (nsp_linear_fwd).

    x: [..., K] (bf16 for prec='bf16', else fp32); w_prepared: tuple from prepare_weight.
    Returns out (and a bf16 copy when out2_bf16).
    """
    _require_cuda(x)
    K = x.shape[-1]
    lead = x.shape[:-1]
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    M = x2.shape[0]
    w = w_prepared[0]
    N = w.shape[0]
    if w.shape[1] != K:                      # weight K was padded at prepare time
        x2 = _pad_k(x2)
        K = x2.shape[1]
        assert K == w.shape[1], (K, w.shape)
    nout = N // 2 if glu else N
    x_lo = None
    if prec == "bf16":
        x2 = to_bf16(x2)
    else:
        x2 = x2.float()
        if prec == "fp32":
            x2, x_lo = split_tf32(x2)
    w_lo = w_prepared[1] if prec == "fp32" else None
    if out is None:
        out = torch.empty(M, nout, dtype=out_dtype, device=x.device)
    out2d = out.reshape(-1, nout)
    res2d = residual.reshape(-1, nout) if residual is not None else None
    out2 = torch.empty(M, nout, dtype=torch.bfloat16, device=x.device) if out2_bf16 else None
    pre = torch.empty(M, N, dtype=torch.bfloat16 if prec == "bf16" else torch.float32, device=x.device) if save_pre else None
    _run("nsp_linear_fwd_save", lib.nsp_linear_fwd_save, PREC[prec], ptr(x2), ptr(x_lo), x2.stride(0), ptr(w), ptr(w_lo), w.stride(0),
         M, N, K, int(glu), ACT[act], ptr(bias), ptr(res2d),
         res2d.stride(0) if res2d is not None else 0, float(alpha),
         ptr(out2d), out2d.stride(0), int(out2d.dtype == torch.bfloat16),
         ptr(out2), out2.stride(0) if out2 is not None else 0, ptr(pre), N, current_stream_ptr(),
         flops=2.0 * M * N * K, tag="gemm_%s" % prec, shape=(M, N, K, "glu" if glu else (act or ""), "res" if res2d is not None else ""))
    res = out2d.reshape(*lead, nout)
    rets = (res,)
    if out2_bf16:
        rets += (out2.reshape(*lead, nout),)
    if save_pre:
        rets += (pre.reshape(*lead, N),)
    return rets if len(rets) > 1 else res


# ---------------------------------------------------------------------------------------------
# normalisation / attention / conformer convolution
# ---------------------------------------------------------------------------------------------
def layernorm(x, weight, bias, eps, out_fp32=True, out_bf16=False, in_scale=1.0):
    """LayerNorm over the last dim of an fp32 tensor (nsp_layernorm_fwd). Returns fp32 and/or bf16 outputs."""
    _require_cuda(x)
    D = x.shape[-1]
    x2 = x.reshape(-1, D)
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    M = x2.shape[0]
    y = torch.empty(M, D, dtype=torch.float32, device=x.device) if out_fp32 else None
    yb = torch.empty(M, D, dtype=torch.bfloat16, device=x.device) if out_bf16 else None
    _run("nsp_layernorm_fwd", lib.nsp_layernorm_fwd, ptr(x2), x2.stride(0), ptr(weight), ptr(bias), float(eps), float(in_scale),
                                ptr(y), D, ptr(yb), D, M, D, current_stream_ptr())
    outs = tuple(t.reshape(x.shape) for t in (y, yb) if t is not None)
    return outs[0] if len(outs) == 1 else outs


def _row_pitch(t, ref_dtype):
    """Row pitch (elements) of a `[B, T, D]` view whose batches lie T rows apart.  torch leaves the strides of size-1
    dimensions arbitrary (a one-frame chunk of a one-utterance batch sliced out of a fused QKV buffer stays a view), so
    they are not trusted."""
    B, T, D = t.shape
    assert t.dtype == ref_dtype and t.stride(2) == 1
    if T > 1:
        assert B == 1 or t.stride(0) == T * t.stride(1)
        return t.stride(1)
    return t.stride(0) if B > 1 else D


def relpos_attention(q, k, v, klens, n_heads, r=None, u_bias=None, v_bias=None, clamp_len=-1, causal=False,
                     lookahead=0, chunk_c=0, chunk_l=0, want_stats=False):
    """Flash-style (rel-pos) self-attention (nsp_relpos_attention_fwd).

    q `[B, Tq, H*dk]`, k/v `[B, Tk, H*dk]` (may be strided views of a fused QKV buffer; last dim contiguous),
    r `[rlen, H*dk]` projected position table or None, klens int32 `[B]` CUDA.  Returns `[B, Tq, H*dk]`.
    """
    _require_cuda(q, k, v, klens)
    B, Tq, D = q.shape
    Tk = k.shape[1]
    dk = D // n_heads
    is_bf16 = q.dtype == torch.bfloat16
    ldq, ldk, ldv = (_row_pitch(t, q.dtype) for t in (q, k, v))
    if r is not None:
        assert r.dtype == q.dtype and r.stride(-1) == 1
        r = r.reshape(-1, D) if r.dim() == 3 else r
    out = torch.empty(B, Tq, D, dtype=q.dtype, device=q.device)
    if want_stats:          # training: keep the softmax statistics for the tensor-core backward (None if the CUDA-core kernel ran)
        import ctypes
        stats = torch.empty(B, n_heads, Tq, 2, dtype=torch.float32, device=q.device)
        written = ctypes.c_int(0)
        _run("nsp_relpos_attention_fwd_stats", lib.nsp_relpos_attention_fwd_stats, int(is_bf16), ptr(q), ldq, ptr(k), ldk,
             ptr(v), ldv, ptr(r), r.stride(0) if r is not None else 0, r.shape[0] if r is not None else 0,
             ptr(u_bias), ptr(v_bias), ptr(klens), ptr(out), D, B, n_heads, Tq, Tk, dk,
             int(clamp_len), int(causal), int(lookahead), int(chunk_c), int(chunk_l), ptr(stats), ctypes.byref(written),
             current_stream_ptr(), flops=4.0 * B * n_heads * Tq * Tk * dk, tag="nsp_relpos_attention_fwd")
        return out, (stats if written.value else None)
    _run("nsp_relpos_attention_fwd", lib.nsp_relpos_attention_fwd, int(is_bf16), ptr(q), ldq, ptr(k), ldk, ptr(v), ldv,
                                       ptr(r), r.stride(0) if r is not None else 0, r.shape[0] if r is not None else 0,
                                       ptr(u_bias), ptr(v_bias), ptr(klens), ptr(out), D, B, n_heads, Tq, Tk, dk,
                                       int(clamp_len), int(causal), int(lookahead), int(chunk_c), int(chunk_l),
                                       current_stream_ptr(), flops=4.0 * B * n_heads * Tq * Tk * dk)
    return out


NORM_MODE = {"layer_norm": 0, "batch_norm": 1, "group_norm": 2}


def conformer_conv(x, dw_weight, dw_bias, norm_mode, norm_w, norm_b, eps, run_mean=None, run_var=None, causal=False):
    """y = Swish(Norm(depthwise_conv(x) + bias)) on `[B, T, d]` (nsp_conformer_conv_fwd)."""
    _require_cuda(x)
    B, T, d = x.shape
    x = x if x.stride(2) == 1 and x.stride(0) == T * x.stride(1) else x.contiguous()
    if dw_weight.dim() == 2:                      # already transposed taps [k, d] (cached by the caller)
        k, w = dw_weight.shape[0], dw_weight
    else:                                         # nn.Conv1d depthwise weight [d, 1, k]
        k = dw_weight.shape[-1]
        w = dw_weight.detach().reshape(d, k).t().contiguous().float()
    y = torch.empty(B, T, d, dtype=x.dtype, device=x.device)
    _run("nsp_conformer_conv_fwd", lib.nsp_conformer_conv_fwd, int(x.dtype == torch.bfloat16), ptr(x), x.stride(1), ptr(w), ptr(dw_bias),
                                     NORM_MODE[norm_mode], ptr(norm_w), ptr(norm_b), ptr(run_mean), ptr(run_var),
                                     float(eps), ptr(y), d, B, T, d, k, int(causal), current_stream_ptr())
    return y


# ---------------------------------------------------------------------------------------------
# small helpers, front-end
# ---------------------------------------------------------------------------------------------
def scale_(x, a):
    _require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    _run("nsp_scale_inplace", lib.nsp_scale_inplace, ptr(x), float(a), x.numel(), current_stream_ptr())
    return x


def mask_rects_(x, freq_rects, time_rects):
    """Zero frequency bands / time spans of x fp32 `[B,T,F]` in place (nsp_mask_rects): lists of (begin, end) pairs."""
    import ctypes
    _require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 3
    B, T, F = x.shape

    def arr(rects):
        flat = [int(v) for r in rects for v in r]
        return (ctypes.c_int32 * max(1, len(flat)))(*flat)
    fa, ta = arr(freq_rects), arr(time_rects)
    _run("nsp_mask_rects", lib.nsp_mask_rects, ptr(x), B, T, F, ctypes.cast(fa, ctypes.c_void_p), len(freq_rects),
         ctypes.cast(ta, ctypes.c_void_p), len(time_rects), current_stream_ptr())
    return x


def add_pos_enc_(x, pe, a=1.0):
    """x[b,t,:] = x[b,t,:] * a + pe[t,:] in place (nsp_add_pos_enc); x fp32 `[B,T,D]` contiguous, pe fp32 `[T,D]`."""
    _require_cuda(x, pe)
    B, T, D = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and pe.dtype == torch.float32 and pe.is_contiguous()
    assert pe.shape == (T, D), (pe.shape, x.shape)
    _run("nsp_add_pos_enc", lib.nsp_add_pos_enc, ptr(x), ptr(pe), float(a), B, T, D, current_stream_ptr())
    return x


def colsum(x):
    _require_cuda(x)
    x = x.float().contiguous()
    M, N = x.shape
    y = torch.empty(N, dtype=torch.float32, device=x.device)
    _run("nsp_colsum", lib.nsp_colsum, ptr(x), ptr(y), M, N, current_stream_ptr())
    return y


def xl_pos_table(inv_freq, rows):
    """[rows, d] TransformerXL sinusoid table, row r = relative distance r (nsp_xl_pos_table)."""
    _require_cuda(inv_freq)
    d = inv_freq.numel() * 2
    tab = torch.empty(rows, d, dtype=torch.float32, device=inv_freq.device)
    _run("nsp_xl_pos_table", lib.nsp_xl_pos_table, ptr(inv_freq), ptr(tab), rows, d, current_stream_ptr())
    return tab


def conv3x3_relu(x, weight, bias, B, T, F, in_chmajor=False, relu=True, out_dtype=torch.float32):
    """Channels-last 3x3 conv + bias + ReLU (nsp_conv3x3_relu_fwd). x holds B*T*F*CI elements."""
    _require_cuda(x)
    CO, CI = weight.shape[0], weight.shape[1]
    assert weight.shape[2:] == (3, 3) and x.numel() == B * T * F * CI and x.is_contiguous()
    y = torch.empty(B, T, F, CO, dtype=out_dtype, device=x.device)
    _run("nsp_conv3x3_relu_fwd", lib.nsp_conv3x3_relu_fwd, int(x.dtype == torch.bfloat16), int(out_dtype == torch.bfloat16), ptr(x),
                                   int(in_chmajor), ptr(weight), ptr(bias), ptr(y), B, T, F, CI, CO, int(relu),
                                   current_stream_ptr())
    return y


def maxpool2d(x, pool_t, pool_f, out_chmajor=False, out_dtype=None):
    """ceil-mode max-pool on channels-last [B,T,F,C] (nsp_maxpool2d_fwd)."""
    _require_cuda(x)
    B, T, F, C = x.shape
    To, Fo = -(-T // pool_t), -(-F // pool_f)
    out_dtype = out_dtype or x.dtype
    shape = (B, To, C * Fo) if out_chmajor else (B, To, Fo, C)
    y = torch.empty(shape, dtype=out_dtype, device=x.device)
    _run("nsp_maxpool2d_fwd", lib.nsp_maxpool2d_fwd, int(x.dtype == torch.bfloat16), int(out_dtype == torch.bfloat16), ptr(x), ptr(y),
                                B, T, F, C, pool_t, pool_f, 0, int(out_chmajor), current_stream_ptr())
    return y


def maxpool_time(x, factor):
    """MaxPoolSubsampler core on [B,T,D] (nsp_maxpool_time_fwd)."""
    _require_cuda(x)
    x = x.contiguous()
    B, T, D = x.shape
    y = torch.empty(B, -(-T // factor), D, dtype=x.dtype, device=x.device)
    _run("nsp_maxpool_time_fwd", lib.nsp_maxpool_time_fwd, int(x.dtype == torch.bfloat16), ptr(x), ptr(y), B, T, D, factor,
                                   current_stream_ptr())
    return y


# ---------------------------------------------------------------------------------------------
# RNN-T
# ---------------------------------------------------------------------------------------------
KERNELS_PER_CALL["nsp_rnnt_loss_fwd_bwd"] = 4


def rnnt_loss_fwd_bwd(log_probs, labels, flens, ylens, blank=0, need_grad=True, return_ws=False):
    """RNN-T loss + d loss/d log_probs (nsp_rnnt_loss_fwd_bwd).  log_probs fp32 `[B,T,U+1,V]` CUDA,
    labels int32 `[B,U]`, flens/ylens int32 `[B]`.  Returns (loss 0-dim mean, nll [B], grad or None)
    [+ the lattice workspace when return_ws: the input of rnnt_grad_logits]."""
    _require_cuda(log_probs, labels, flens, ylens)
    log_probs = log_probs.contiguous()
    assert log_probs.dtype == torch.float32 and log_probs.dim() == 4
    B, T, U1, V = log_probs.shape
    assert labels.shape == (B, U1 - 1) and labels.is_contiguous()
    ws_bytes = lib.nsp_rnnt_loss_workspace_bytes(B, T, U1)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=log_probs.device)
    nll = torch.empty(B, dtype=torch.float32, device=log_probs.device)
    loss = torch.empty((), dtype=torch.float32, device=log_probs.device)
    grad = torch.empty_like(log_probs) if need_grad else None
    _run("nsp_rnnt_loss_fwd_bwd", lib.nsp_rnnt_loss_fwd_bwd, ptr(log_probs), B, T, U1, V, ptr(labels), ptr(flens),
         ptr(ylens), int(blank), ptr(nll), ptr(loss), ptr(grad), ptr(ws), ws_bytes, current_stream_ptr(),
         nbytes=8.0 * B * T * U1 * V, tag="rnnt_loss")
    if return_ws:
        return loss, nll, grad, ws
    return loss, nll, grad


def rnnt_grad_logits(log_probs, ws, nll, labels, flens, ylens, blank=0, gscale=None, out_dtype=torch.float32, inplace=False):
    """d loss / d logits in one pass (nsp_rnnt_grad_logits) from log_probs = log_softmax(logits) and the lattice
    workspace of rnnt_loss_fwd_bwd(..., need_grad=False, return_ws=True).  fp32 (optionally in place over log_probs)
    or bf16 `[B,T,U+1,V]`."""
    _require_cuda(log_probs, ws, nll, labels, flens, ylens, gscale)
    assert log_probs.dtype == torch.float32 and log_probs.is_contiguous() and log_probs.dim() == 4
    B, T, U1, V = log_probs.shape
    if gscale is not None:
        gscale = gscale.reshape(-1)[:1].float().contiguous()
    if inplace and out_dtype == torch.float32:
        dz = log_probs
    else:
        dz = torch.empty(B, T, U1, V, dtype=out_dtype, device=log_probs.device)
    _run("nsp_rnnt_grad_logits", lib.nsp_rnnt_grad_logits, ptr(log_probs), B, T, U1, V, ptr(labels), ptr(flens), ptr(ylens),
         int(blank), ptr(nll), ptr(ws), ws.numel(), ptr(gscale), ptr(dz), int(out_dtype == torch.bfloat16),
         current_stream_ptr(), nbytes=float(log_probs.numel() * (4 + dz.element_size())), tag="rnnt_grad_logits")
    return dz


def conv3x3_c32_tc(x, w_taps, bias, relu=True, pool2x2=False):
    """tcgen05 implicit-GEMM 3x3 conv for 32->32 channels (nsp_conv3x3_c32_tc_fwd).
    x bf16 `[B,T,F,32]`, w_taps bf16 `[32,288]` (column = tap*32 + ci)."""
    _require_cuda(x, w_taps)
    B, T, F, C = x.shape
    assert C == 32 and x.dtype == torch.bfloat16 and x.is_contiguous() and w_taps.shape == (32, 288)
    To, Fo = (-(-T // 2), -(-F // 2)) if pool2x2 else (T, F)
    y = torch.empty(B, To, Fo, 32, dtype=torch.bfloat16, device=x.device)
    _run("nsp_conv3x3_c32_tc_fwd", lib.nsp_conv3x3_c32_tc_fwd, ptr(x), ptr(w_taps), ptr(bias), ptr(y), B, T, F,
         int(relu), int(pool2x2), current_stream_ptr(), flops=2.0 * B * T * F * 32 * 288, tag="conv3x3_tc")
    return y


# ---------------------------------------------------------------------------------------------
# softmax rows, greedy CTC, RNN-T joint
# ---------------------------------------------------------------------------------------------
KERNELS_PER_CALL["nsp_ctc_greedy"] = 2


def softmax_rows(x, log=False, temperature=1.0, inplace=False):
    """(log-)softmax over the last dim of an fp32 tensor (nsp_softmax_rows)."""
    _require_cuda(x)
    x = x.contiguous()
    assert x.dtype == torch.float32
    V = x.shape[-1]
    y = x if inplace else torch.empty_like(x)
    _run("nsp_softmax_rows", lib.nsp_softmax_rows, ptr(x), ptr(y), x.numel() // V, V, int(log), float(temperature),
         current_stream_ptr(), nbytes=8.0 * x.numel())
    return y


def ctc_greedy(logits, elens, blank=0):
    """Greedy CTC path on the device (nsp_ctc_greedy) -> (best [B,T], hyp [B,T], hyp_lens [B], trigger [B,T]) int32."""
    _require_cuda(logits, elens)
    logits = logits.contiguous().float()
    B, T, V = logits.shape
    mk = lambda *s: torch.zeros(*s, dtype=torch.int32, device=logits.device)   # noqa: E731
    best, hyp, hyp_lens, trig = mk(B, T), mk(B, T), mk(B), mk(B, T)
    _run("nsp_ctc_greedy", lib.nsp_ctc_greedy, ptr(logits), B, T, V, ptr(elens), int(blank), ptr(best), ptr(hyp),
         ptr(hyp_lens), ptr(trig), current_stream_ptr())
    return best, hyp, hyp_lens, trig


def rnnt_joint_tanh(enc, dec, out_dtype=torch.float32):
    """tanh(enc[:, :, None] + dec[:, None]) (nsp_rnnt_joint_tanh): enc `[B,T,J]`, dec `[B,U1,J]` fp32 -> `[B,T,U1,J]`."""
    _require_cuda(enc, dec)
    enc, dec = enc.contiguous().float(), dec.contiguous().float()
    B, T, J = enc.shape
    U1 = dec.shape[1]
    out = torch.empty(B, T, U1, J, dtype=out_dtype, device=enc.device)
    _run("nsp_rnnt_joint_tanh", lib.nsp_rnnt_joint_tanh, ptr(enc), ptr(dec), ptr(out), int(out_dtype == torch.bfloat16),
         B, T, U1, J, current_stream_ptr())
    return out


POOL_MODE = {"max": 0, "mean": 1, "drop": 2, "add": 3}


def pool_time(x, factor, mode):
    """Time pooling on `[B,T,D]` with kernel = stride = factor, ceil-mode (nsp_pool_time_fwd)."""
    _require_cuda(x)
    x = x.contiguous()
    B, T, D = x.shape
    y = torch.empty(B, -(-T // factor), D, dtype=x.dtype, device=x.device)
    _run("nsp_pool_time_fwd", lib.nsp_pool_time_fwd, int(x.dtype == torch.bfloat16), ptr(x), ptr(y), B, T, D, factor,
         POOL_MODE[mode], current_stream_ptr())
    return y


def _lstm_tc(prec, B, H, n_dirs):
    """bf16 mode takes the tensor-core recurrence (lstm_tc.cu) when the shape is covered; NSP_LSTM_PATH=simt forces lstm.cu."""
    return prec == "bf16" and os.environ.get("NSP_LSTM_PATH", "") != "simt" and bool(lib.nsp_lstm_tc_supported(B, H, n_dirs))


def lstm_seq(gates_x, w_hh, lens, n_dirs, save=False, state=None, want_state=False, prec=None):
    """LSTM recurrence of one layer (nsp_lstm_seq_fwd): gates_x fp32 `[B,T,n_dirs*4H]`, w_hh fp32 `[n_dirs,4H,H]`,
    lens int32 `[B]` CUDA -> y fp32 `[B,T,n_dirs*H]` (zeros beyond each length).
    save=True (training, nsp_lstm_seq_fwd_save) -> (y, acts `[B,T,n_dirs,4H]`, cprev, hprev `[B,T,n_dirs,H]`).
    state=(h0, c0) fp32 `[n_dirs,B,H]` (nn.LSTM's hx) / want_state=True (streaming, nsp_lstm_seq_fwd_state)
    -> (y, (hN, cN)).  prec="bf16": the recurrent product runs on the tensor cores with bf16 operands
    (nsp_lstm_seq_fwd_tc; cell state, gate math and every output stay fp32); otherwise fp32 CUDA-core math."""
    _require_cuda(gates_x, w_hh, lens)
    gates_x = gates_x.contiguous().float()
    w_hh = w_hh.contiguous().float()
    B, T, G = gates_x.shape
    H = G // (4 * n_dirs)
    if _lstm_tc(prec, B, H, n_dirs):
        dev = gates_x.device
        ws_bytes = lib.nsp_lstm_tc_workspace_bytes(B, H, n_dirs, 0)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        y = torch.empty(B, T, n_dirs * H, dtype=torch.float32, device=dev)
        h0 = c0 = hN = cN = acts = cprev = hprev = None
        if state is not None:
            h0, c0 = (t.contiguous().float() for t in state)
            _require_cuda(h0, c0)
            assert h0.shape == (n_dirs, B, H) and c0.shape == (n_dirs, B, H), (h0.shape, c0.shape, (n_dirs, B, H))
        if state is not None or want_state:
            hN = torch.empty(n_dirs, B, H, dtype=torch.float32, device=dev)
            cN = torch.empty(n_dirs, B, H, dtype=torch.float32, device=dev)
        if save:
            acts = torch.zeros(B, T, n_dirs, 4 * H, dtype=torch.float32, device=dev)
            cprev = torch.zeros(B, T, n_dirs, H, dtype=torch.float32, device=dev)
            hprev = torch.zeros(B, T, n_dirs, H, dtype=torch.float32, device=dev)
        _run("nsp_lstm_seq_fwd_tc", lib.nsp_lstm_seq_fwd_tc, ptr(gates_x), ptr(w_hh), ptr(lens), ptr(y), B, T, H, n_dirs,
             ptr(acts), ptr(cprev), ptr(hprev), ptr(h0), ptr(c0), ptr(hN), ptr(cN), ptr(ws), ws_bytes, current_stream_ptr(),
             flops=2.0 * B * T * n_dirs * 4 * H * H, tag="lstm_seq")
        out = (y, acts, cprev, hprev) if save else (y,)
        if hN is not None:
            out = out + ((hN, cN),)
        return out if len(out) > 1 else out[0]
    ws_bytes = lib.nsp_lstm_workspace_bytes(B, H, n_dirs)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=gates_x.device)
    y = torch.empty(B, T, n_dirs * H, dtype=torch.float32, device=gates_x.device)
    h0 = c0 = hN = cN = None
    if state is not None or want_state:
        if state is not None:
            h0, c0 = (t.contiguous().float() for t in state)
            _require_cuda(h0, c0)
            assert h0.shape == (n_dirs, B, H) and c0.shape == (n_dirs, B, H), (h0.shape, c0.shape, (n_dirs, B, H))
        hN = torch.empty(n_dirs, B, H, dtype=torch.float32, device=gates_x.device)
        cN = torch.empty(n_dirs, B, H, dtype=torch.float32, device=gates_x.device)
        if not save:
            _run("nsp_lstm_seq_fwd_state", lib.nsp_lstm_seq_fwd_state, ptr(gates_x), ptr(w_hh), ptr(lens), ptr(y), B, T, H, n_dirs,
                 ptr(h0), ptr(c0), ptr(hN), ptr(cN), ptr(ws), ws_bytes, current_stream_ptr(),
                 flops=2.0 * B * T * n_dirs * 4 * H * H, tag="lstm_seq")
            return y, (hN, cN)
    if save:
        acts = torch.zeros(B, T, n_dirs, 4 * H, dtype=torch.float32, device=gates_x.device)
        cprev = torch.zeros(B, T, n_dirs, H, dtype=torch.float32, device=gates_x.device)
        hprev = torch.zeros(B, T, n_dirs, H, dtype=torch.float32, device=gates_x.device)
        if hN is not None:          # training with a carried state (latency-controlled BLSTM)
            _run("nsp_lstm_seq_fwd_save_state", lib.nsp_lstm_seq_fwd_save_state, ptr(gates_x), ptr(w_hh), ptr(lens), ptr(y), B, T, H,
                 n_dirs, ptr(acts), ptr(cprev), ptr(hprev), ptr(h0), ptr(c0), ptr(hN), ptr(cN), ptr(ws), ws_bytes,
                 current_stream_ptr(), flops=2.0 * B * T * n_dirs * 4 * H * H, tag="lstm_seq")
            return y, acts, cprev, hprev, (hN, cN)
        _run("nsp_lstm_seq_fwd_save", lib.nsp_lstm_seq_fwd_save, ptr(gates_x), ptr(w_hh), ptr(lens), ptr(y), B, T, H, n_dirs,
             ptr(acts), ptr(cprev), ptr(hprev), ptr(ws), ws_bytes, current_stream_ptr(),
             flops=2.0 * B * T * n_dirs * 4 * H * H, tag="lstm_seq")
        return y, acts, cprev, hprev
    _run("nsp_lstm_seq_fwd", lib.nsp_lstm_seq_fwd, ptr(gates_x), ptr(w_hh), ptr(lens), ptr(y), B, T, H, n_dirs, ptr(ws),
         ws_bytes, current_stream_ptr(), flops=2.0 * B * T * n_dirs * 4 * H * H, tag="lstm_seq")
    return y


def lstm_seq_bwd(dy, acts, cprev, w_hh, lens, dstate=None, want_dstate=False, prec=None):
    """Backpropagation through time of lstm_seq (nsp_lstm_seq_bwd): dy fp32 `[B,T,n_dirs*H]` + what the forward saved ->
    d loss / d gate pre-activations fp32 `[B,T,n_dirs*4H]` (the layout of gates_x; zero beyond each length).
    dstate = (dhN, dcN): gradient w.r.t. the final state; want_dstate -> (dg, (dh0, dc0)) (nsp_lstm_seq_bwd_state).
    prec="bf16": dG_t W_hh on the tensor cores (nsp_lstm_seq_bwd_tc)."""
    _require_cuda(dy, acts, cprev, w_hh, lens)
    B, T, n_dirs, H4 = acts.shape
    H = H4 // 4
    dy = dy.contiguous().float()
    assert dy.shape == (B, T, n_dirs * H), (dy.shape, acts.shape)
    w_hh = w_hh.contiguous().float()
    dg = torch.empty(B, T, n_dirs * H4, dtype=torch.float32, device=dy.device)
    if _lstm_tc(prec, B, H, n_dirs):
        ws_bytes = lib.nsp_lstm_tc_workspace_bytes(B, H, n_dirs, 1)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dy.device)
        dhN = dcN = dh0 = dc0 = None
        if dstate is not None:
            dhN, dcN = (t.contiguous().float() for t in dstate)
            assert dhN.shape == (n_dirs, B, H) and dcN.shape == (n_dirs, B, H)
        if want_dstate:
            dh0 = torch.zeros(n_dirs, B, H, dtype=torch.float32, device=dy.device)
            dc0 = torch.zeros(n_dirs, B, H, dtype=torch.float32, device=dy.device)
        _run("nsp_lstm_seq_bwd_tc", lib.nsp_lstm_seq_bwd_tc, ptr(dy), ptr(acts), ptr(cprev), ptr(w_hh), ptr(lens), ptr(dg),
             B, T, H, n_dirs, ptr(dhN), ptr(dcN), ptr(dh0), ptr(dc0), ptr(ws), ws_bytes, current_stream_ptr(),
             flops=2.0 * B * T * n_dirs * 4 * H * H, tag="lstm_seq_bwd")
        return (dg, (dh0, dc0)) if want_dstate else dg
    ws = torch.empty(256, dtype=torch.uint8, device=dy.device)
    if dstate is not None or want_dstate:
        dhN = dcN = dh0 = dc0 = None
        if dstate is not None:
            dhN, dcN = (t.contiguous().float() for t in dstate)
            assert dhN.shape == (n_dirs, B, H) and dcN.shape == (n_dirs, B, H)
        if want_dstate:
            dh0 = torch.zeros(n_dirs, B, H, dtype=torch.float32, device=dy.device)
            dc0 = torch.zeros(n_dirs, B, H, dtype=torch.float32, device=dy.device)
        _run("nsp_lstm_seq_bwd_state", lib.nsp_lstm_seq_bwd_state, ptr(dy), ptr(acts), ptr(cprev), ptr(w_hh), ptr(lens), ptr(dg),
             B, T, H, n_dirs, ptr(dhN), ptr(dcN), ptr(dh0), ptr(dc0), ptr(ws), 256, current_stream_ptr(),
             flops=2.0 * B * T * n_dirs * 4 * H * H, tag="lstm_seq_bwd")
        return (dg, (dh0, dc0)) if want_dstate else dg
    _run("nsp_lstm_seq_bwd", lib.nsp_lstm_seq_bwd, ptr(dy), ptr(acts), ptr(cprev), ptr(w_hh), ptr(lens), ptr(dg), B, T, H, n_dirs,
         ptr(ws), 256, current_stream_ptr(), flops=2.0 * B * T * n_dirs * 4 * H * H, tag="lstm_seq_bwd")
    return dg


# ---------------------------------------------------------------------------------------------
# backward pass (training): hand-written gradients of the forward ops above
# ---------------------------------------------------------------------------------------------
KERNELS_PER_CALL["nsp_relpos_attention_bwd"] = 4
KERNELS_PER_CALL["nsp_rnnt_joint_tanh_bwd"] = 2
KERNELS_PER_CALL["nsp_bn_swish_bwd"] = 2
KERNELS_PER_CALL["nsp_conformer_conv_bwd"] = 3


def _operand(x, prec):
    """2-D contiguous operand in the GEMM dtype of `prec`: bf16, fp32, or (hi, lo) for 'fp32'."""
    x = x.reshape(-1, x.shape[-1])
    if prec == "bf16":
        return to_bf16(x.contiguous()), None
    x = x.float().contiguous()
    if prec == "fp32":
        return split_tf32(x)
    return x, None


def linear_wgrad(dy, x, prec, dw, alpha=1.0, accumulate=True, side=False):
    """dw[N,K] (+)= alpha * dy[M,N]^T x[M,K] on the tcgen05 MN-major wgrad kernel (nsp_linear_wgrad)."""
    _require_cuda(dy, x, dw)
    assert dw.dtype == torch.float32 and dw.dim() == 2 and dw.stride(1) == 1
    N, K = dw.shape
    dy2, x2 = dy.reshape(-1, N), x.reshape(-1, x.shape[-1])[:, :K]
    M = dy2.shape[0]
    assert x2.shape[0] == M, (dy.shape, x.shape)
    mult = 8 if prec == "bf16" else 4
    if N % mult or K % mult:                       # TMA pitches are 16-byte multiples: pad odd widths through a temporary
        Np, Kp = -(-N // mult) * mult, -(-K // mult) * mult
        tmp = torch.zeros(Np, Kp, dtype=torch.float32, device=dw.device)
        linear_wgrad(torch.nn.functional.pad(dy2, (0, Np - N)), torch.nn.functional.pad(x2, (0, Kp - K)), prec, tmp, alpha, True)
        if accumulate:
            dw.add_(tmp[:N, :K])
        else:
            dw.copy_(tmp[:N, :K])
        return dw
    if prec != "bf16":
        # parity modes: tcgen05 has no 128B-swizzled MN-major layout for tf32 operands, so the fp32 / tf32 weight gradient
        # runs on the K-major GEMM (3xTF32 in 'fp32' mode) over transposed copies (pure data movement; not the perf path)
        a_t = dy2.float().t().contiguous()
        b_t = prepare_weight(x2.float().t().contiguous(), prec)
        linear(a_t, b_t, None, prec=prec, residual=dw if accumulate else None, alpha=alpha, out_dtype=torch.float32, out=dw)
        return dw
    dyh, dyl = _operand(dy2, prec)
    xh, xl = _operand(x2, prec)

    def launch():
        _run("nsp_linear_wgrad", lib.nsp_linear_wgrad, PREC[prec], ptr(dyh), ptr(dyl), dyh.stride(0), ptr(xh), ptr(xl), xh.stride(0),
             M, N, K, float(alpha), ptr(dw), dw.stride(0), int(accumulate), current_stream_ptr(),
             flops=2.0 * M * N * K, tag="gemm_wgrad_%s" % prec, shape=(M, N, K))

    if side and wgrad_side_enabled() and torch.cuda.is_available():
        # Weight gradients feed nothing until the optimizer / all-reduce: run them on a second stream next to the input-gradient
        # chain (they fill the SMs that chain leaves idle in its launch tails).  Operands stay referenced until wgrad_join().
        cur = torch.cuda.current_stream()
        st = _WG["streams"].get(dw.device)
        if st is None:
            st = _WG["streams"][dw.device] = torch.cuda.Stream(device=dw.device)
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            launch()
        _WG["keep"].extend((dyh, dyl, xh, xl, dw))
        _WG["pending"].add(st)
        if not _WG["cb"]:
            try:                                   # once per backward pass: join when the autograd engine is done
                torch.autograd.Variable._execution_engine.queue_callback(wgrad_join)
                _WG["cb"] = True
            except RuntimeError:                   # not inside a backward pass: the caller joins
                pass
        return dw
    launch()
    return dw


_WG = {"streams": {}, "keep": [], "pending": set(), "cb": False}


def wgrad_side_enabled():
    """Weight-gradient GEMMs of the encoder blocks go to a side stream (see linear_wgrad): 19.36 -> 18.71 ms per Conformer-L
    training step on B200 (A/B in one call, profiles/README.md round 2).  NSP_WGRAD_STREAM=0 keeps them on the caller's stream."""
    return os.environ.get("NSP_WGRAD_STREAM", "1") != "0"


def wgrad_join():
    """The current stream waits for every weight gradient issued on the side stream (end of backward, or before a node's
    gradient bucket is handed to the all-reduce hook)."""
    if _WG["pending"]:
        cur = torch.cuda.current_stream()
        for st in _WG["pending"]:
            cur.wait_stream(st)
        _WG["pending"].clear()
    del _WG["keep"][:]
    _WG["cb"] = False


def layernorm_bwd(dy, x, gamma, eps, dres=None, dgamma=None, dbeta=None, want_fp32=True, want_bf16=False, in_scale=1.0,
                  dcol=None, dcol_alpha=1.0):
    """dx = dres + LN'(x).dy (nsp_layernorm_bwd); dgamma/dbeta fp32 [D] are accumulated.  Returns fp32 and/or bf16 dx."""
    _require_cuda(dy, x)
    D = x.shape[-1]
    dy2, x2 = dy.reshape(-1, D).float(), x.reshape(-1, D)
    dy2 = dy2 if dy2.stride(1) == 1 else dy2.contiguous()
    M = x2.shape[0]
    dr2 = dres.reshape(-1, D) if dres is not None else None
    dx = torch.empty(M, D, dtype=torch.float32, device=x.device) if want_fp32 else None
    dxb = torch.empty(M, D, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    _run("nsp_layernorm_bwd", lib.nsp_layernorm_bwd, ptr(dy2), dy2.stride(0), ptr(x2), x2.stride(0), ptr(gamma), float(eps),
         float(in_scale), ptr(dr2), dr2.stride(0) if dr2 is not None else 0, ptr(dx), D, ptr(dxb), D,
         ptr(dgamma), ptr(dbeta), ptr(dcol), float(dcol_alpha), M, D, current_stream_ptr(), nbytes=M * D * 12.0)
    outs = tuple(t.reshape(x.shape) for t in (dx, dxb) if t is not None)
    return outs[0] if len(outs) == 1 else outs


def act_bwd(dh, z, act):
    """dz = dh * act'(z) (nsp_act_bwd)."""
    _require_cuda(dh, z)
    assert dh.dtype == z.dtype and dh.shape == z.shape
    dh, z = dh.contiguous(), z.contiguous()
    dz = torch.empty_like(dh)
    _run("nsp_act_bwd", lib.nsp_act_bwd, int(dh.dtype == torch.bfloat16), ACT[act], ptr(dh), ptr(z), ptr(dz), dh.numel(),
         current_stream_ptr())
    return dz


def act_bwd_bias(dh, z, act, dbias, glu=False):
    """act_bwd / glu_bwd with the bias gradient fused (nsp_act_bwd_bias): dbias fp32 [N] += column sums of the result.
    bf16, vectorisable widths only -- callers fall back to act_bwd / glu_bwd + colsum_acc otherwise."""
    _require_cuda(dh, z, dbias)
    N = z.shape[-1]
    dh, z = dh.contiguous(), z.contiguous()
    dz = torch.empty_like(z)
    _run("nsp_act_bwd_bias", lib.nsp_act_bwd_bias, int(glu), 0 if glu else ACT[act], ptr(dh), ptr(z), ptr(dz), ptr(dbias),
         z.numel() // N, N, current_stream_ptr())
    return dz


def glu_bwd(dg, pre):
    """Backward of out = a * sigmoid(b), pre = [a | b] (nsp_glu_bwd) -> dpre like pre."""
    _require_cuda(dg, pre)
    d = dg.shape[-1]
    dg, pre = dg.contiguous(), pre.contiguous()
    assert pre.shape[-1] == 2 * d and dg.dtype == pre.dtype
    dpre = torch.empty_like(pre)
    _run("nsp_glu_bwd", lib.nsp_glu_bwd, int(dg.dtype == torch.bfloat16), ptr(dg), ptr(pre), ptr(dpre), dg.numel() // d, d,
         current_stream_ptr())
    return dpre


def colsum_acc(x, y, alpha=1.0):
    """y[N] += alpha * column sums of x [M,N] (bf16 or fp32) (nsp_colsum_acc)."""
    _require_cuda(x, y)
    x2 = x.reshape(-1, x.shape[-1])
    x2 = x2 if x2.stride(1) == 1 else x2.contiguous()
    if x2.dtype not in (torch.bfloat16, torch.float32):
        x2 = x2.float()
    M, N = x2.shape
    if N < 128 and 256 % N == 0 and x2.is_contiguous() and M % (256 // N) == 0:
        # narrow matrix (e.g. 32 channels): view as [M*N/256, 256] so that every lane of the kernel has work, then fold the
        # 256 partial columns back onto the N real ones
        tmp = torch.zeros(256, dtype=torch.float32, device=x2.device)
        xw = x2.view(-1, 256)
        _run("nsp_colsum_acc", lib.nsp_colsum_acc, int(xw.dtype == torch.bfloat16), ptr(xw), 256, xw.shape[0], 256, float(alpha),
             ptr(tmp), current_stream_ptr())
        part = tmp.view(256 // N, N)
        _run("nsp_colsum_acc", lib.nsp_colsum_acc, 0, ptr(part), N, 256 // N, N, 1.0, ptr(y), current_stream_ptr())
        return y
    _run("nsp_colsum_acc", lib.nsp_colsum_acc, int(x2.dtype == torch.bfloat16), ptr(x2), x2.stride(0), M, N,
         float(alpha), ptr(y), current_stream_ptr())
    return y


def maxpool_time_bwd(x, dy, factor):
    """MaxPoolSubsampler backward (nsp_maxpool_time_bwd): x fp32 [B,T,D] forward input, dy fp32 [B,T',D]."""
    _require_cuda(x, dy)
    x, dy = x.contiguous().float(), dy.contiguous().float()
    B, T, D = x.shape
    dx = torch.empty_like(x)
    _run("nsp_maxpool_time_bwd", lib.nsp_maxpool_time_bwd, ptr(x), ptr(dy), ptr(dx), B, T, D, int(factor), current_stream_ptr())
    return dx


def pool_time_bwd(dy, T, factor, mode):
    """Backward of pool_time for mode 'mean' / 'drop' / 'add' (nsp_pool_time_bwd): dy fp32 `[B,ceil(T/f),D]` -> dx `[B,T,D]`."""
    _require_cuda(dy)
    dy = dy.contiguous().float()
    B, To, D = dy.shape
    assert To == -(-T // factor) and mode in ("mean", "drop", "add")
    dx = torch.empty(B, T, D, dtype=torch.float32, device=dy.device)
    _run("nsp_pool_time_bwd", lib.nsp_pool_time_bwd, ptr(dy), ptr(dx), B, T, D, int(factor), POOL_MODE[mode], current_stream_ptr())
    return dx


def dwconv_stats(x, taps, dw_bias, causal=False):
    """z = depthwise_conv(x) + bias `[B,T,d]` (dtype of x) and its per-channel (sum, sum of squares) over all B*T frames,
    fp32 `[2,d]` (nsp_dwconv_stats_fwd): the first half of the BatchNorm training forward of the Conformer conv module."""
    _require_cuda(x, taps, dw_bias)
    B, T, d = x.shape
    x = x if x.stride(2) == 1 and x.stride(0) == T * x.stride(1) else x.contiguous()
    z = torch.empty(B, T, d, dtype=x.dtype, device=x.device)
    stats = torch.empty(2, d, dtype=torch.float32, device=x.device)
    _run("nsp_dwconv_stats_fwd", lib.nsp_dwconv_stats_fwd, int(x.dtype == torch.bfloat16), ptr(x), x.stride(1), ptr(taps),
         ptr(dw_bias), ptr(z), d, ptr(stats), B, T, d, taps.shape[0], int(causal), current_stream_ptr())
    return z, stats


def bn_swish_bwd(z, dy, mean, var, gamma, beta, eps):
    """Backward of Swish(BatchNorm(z)) with batch statistics (nsp_bn_swish_bwd): -> (dz like z, sums fp32 `[2,d]` =
    (d beta, d gamma))."""
    _require_cuda(z, dy, mean, var)
    B, T, d = z.shape
    z = z.contiguous()
    dy = dy.to(z.dtype).contiguous()
    dz = torch.empty_like(z)
    sums = torch.empty(2, d, dtype=torch.float32, device=z.device)
    _run("nsp_bn_swish_bwd", lib.nsp_bn_swish_bwd, int(z.dtype == torch.bfloat16), ptr(z), d, ptr(dy), d, ptr(mean), ptr(var),
         ptr(gamma), ptr(beta), float(eps), ptr(sums), ptr(dz), d, B * T, d, current_stream_ptr())
    return dz, sums


def bn_bwd(z, du, mean, var, gamma, eps):
    """BatchNorm backward with batch statistics, no activation inside (nsp_bn_bwd): z, du `[M, d]` (same dtype) ->
    (dz like z, sums fp32 `[2, d]` = (d beta, d gamma))."""
    _require_cuda(z, du, mean, var, gamma)
    M, d = z.shape
    z = z.contiguous()
    du = du.to(z.dtype).contiguous()
    assert du.shape == z.shape and mean.dtype == var.dtype == gamma.dtype == torch.float32
    dz = torch.empty_like(z)
    sums = torch.empty(2, d, dtype=torch.float32, device=z.device)
    _run("nsp_bn_bwd", lib.nsp_bn_bwd, int(z.dtype == torch.bfloat16), ptr(z), d, ptr(du), d, ptr(mean), ptr(var), ptr(gamma),
         float(eps), ptr(sums), ptr(dz), d, M, d, current_stream_ptr(), nbytes=M * d * 3.0 * z.element_size())
    return dz, sums


def gn2_swish_bwd(z, dy, gamma, beta, eps, dgamma, dbeta):
    """Backward of Swish(GroupNorm(z)) with 2 channels per group on the per-frame view (nsp_gn2_swish_bwd): -> dz like z;
    dgamma / dbeta fp32 `[d]` are accumulated."""
    _require_cuda(z, dy, dgamma, dbeta)
    B, T, d = z.shape
    z = z.contiguous()
    dy = dy.to(z.dtype).contiguous()
    dz = torch.empty_like(z)
    _run("nsp_gn2_swish_bwd", lib.nsp_gn2_swish_bwd, int(z.dtype == torch.bfloat16), ptr(z), d, ptr(dy), d, ptr(gamma), ptr(beta),
         float(eps), ptr(dz), d, ptr(dgamma), ptr(dbeta), B * T, d, current_stream_ptr())
    return dz


def dwconv_bwd(x, taps, dz, dtaps, dbias, causal=False):
    """dx, d taps (+=), d bias (+=) of the depthwise conv from dz (nsp_dwconv_bwd)."""
    _require_cuda(x, taps, dz)
    B, T, d = x.shape
    x = x if x.stride(2) == 1 and x.stride(0) == T * x.stride(1) else x.contiguous()
    dz = dz.to(x.dtype).contiguous()
    dx = torch.empty(B, T, d, dtype=x.dtype, device=x.device)
    _run("nsp_dwconv_bwd", lib.nsp_dwconv_bwd, int(x.dtype == torch.bfloat16), ptr(x), x.stride(1), ptr(taps), ptr(dz), d,
         ptr(dx), d, ptr(dtaps), ptr(dbias), B, T, d, taps.shape[0], int(causal), current_stream_ptr())
    return dx


def log_softmax_bwd_(lp, dlp, gscale=None):
    """dz = g * (dlp - exp(lp) * rowsum(dlp)) in place on dlp (nsp_log_softmax_bwd); lp, dlp fp32 `[..., V]` contiguous,
    gscale = optional 0-dim / 1-element fp32 CUDA tensor (upstream gradient of the loss)."""
    _require_cuda(lp, dlp, gscale)
    assert lp.dtype == torch.float32 and dlp.dtype == torch.float32 and lp.is_contiguous() and dlp.is_contiguous()
    assert lp.shape == dlp.shape
    V = lp.shape[-1]
    if gscale is not None:
        gscale = gscale.reshape(-1)[:1].float().contiguous()
    _run("nsp_log_softmax_bwd", lib.nsp_log_softmax_bwd, ptr(lp), ptr(dlp), lp.numel() // V, V, ptr(gscale), current_stream_ptr(),
         nbytes=12.0 * lp.numel())
    return dlp


def rnnt_joint_tanh_bwd(h, dh):
    """Backward of rnnt_joint_tanh (nsp_rnnt_joint_tanh_bwd): h, dh `[B,T,U1,J]` (same dtype, fp32 or bf16) ->
    (de fp32 `[B,T,J]`, dd fp32 `[B,U1,J]`)."""
    _require_cuda(h, dh)
    h, dh = h.contiguous(), dh.contiguous()
    assert h.dtype == dh.dtype and h.shape == dh.shape and h.dim() == 4
    B, T, U1, J = h.shape
    de = torch.empty(B, T, J, dtype=torch.float32, device=h.device)
    dd = torch.empty(B, U1, J, dtype=torch.float32, device=h.device)
    _run("nsp_rnnt_joint_tanh_bwd", lib.nsp_rnnt_joint_tanh_bwd, int(h.dtype == torch.bfloat16), ptr(h), ptr(dh), ptr(de), ptr(dd),
         B, T, U1, J, current_stream_ptr(), nbytes=4.0 * h.numel() * h.element_size())
    return de, dd


# ---- dropout (mask regenerated from (seed, offset, stream, index); nothing is stored) ----
def dropout(x, p, stream_id, scale=1.0, out_dtype=None, inplace=False):
    """y = keep ? x * scale / (1 - p) : 0 (nsp_dropout).  x fp32 / bf16 contiguous; the same (p, stream_id) applied to a
    gradient of the same shape reproduces the mask (backward)."""
    from . import random as nrandom
    _require_cuda(x)
    x = x if x.is_contiguous() else x.contiguous()
    out_dtype = out_dtype or x.dtype
    assert x.dtype in (torch.float32, torch.bfloat16) and out_dtype in (torch.float32, torch.bfloat16)
    y = x if (inplace and out_dtype == x.dtype) else torch.empty(x.shape, dtype=out_dtype, device=x.device)
    _run("nsp_dropout", lib.nsp_dropout, int(x.dtype == torch.bfloat16), int(out_dtype == torch.bfloat16), ptr(x), ptr(y),
         x.numel(), float(p), float(scale), ptr(nrandom.state(x.device)), int(stream_id), current_stream_ptr(),
         nbytes=float(x.numel() * (x.element_size() + y.element_size())))
    return y


def dropout_add(t, res, p, alpha, stream_id, out=None):
    """out = res + alpha * dropout(t) (nsp_dropout_add): residual-branch output; res / out fp32, t fp32 or bf16."""
    from . import random as nrandom
    _require_cuda(t, res)
    t = t if t.is_contiguous() else t.contiguous()
    assert res.dtype == torch.float32 and res.is_contiguous() and res.numel() == t.numel()
    out = torch.empty_like(res) if out is None else out
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == t.numel()
    _run("nsp_dropout_add", lib.nsp_dropout_add, int(t.dtype == torch.bfloat16), ptr(t), ptr(res), ptr(out), t.numel(),
         float(p), float(alpha), ptr(nrandom.state(t.device)), int(stream_id), current_stream_ptr(),
         nbytes=float(t.numel() * (t.element_size() + 8)))
    return out


def rng_advance(state):
    _require_cuda(state)
    _run("nsp_rng_advance", lib.nsp_rng_advance, ptr(state), current_stream_ptr())


def relu_mask(dx, a):
    """dz = a > 0 ? dx : 0 (nsp_relu_mask)."""
    _require_cuda(dx, a)
    dx, a = dx.contiguous(), a.contiguous()
    assert dx.dtype == a.dtype and dx.numel() == a.numel()
    dz = torch.empty_like(dx)
    _run("nsp_relu_mask", lib.nsp_relu_mask, int(dx.dtype == torch.bfloat16), ptr(dx), ptr(a), ptr(dz), dx.numel(), current_stream_ptr())
    return dz


def maxpool2d_relu_bwd(a, dy, pool_t, pool_f, in_chmajor=False):
    """ReLU + ceil-mode max-pool backward on channels-last a `[B,T,F,C]` (nsp_maxpool2d_relu_bwd)."""
    _require_cuda(a, dy)
    a, dy = a.contiguous(), dy.contiguous()
    B, T, F, C = a.shape
    dz = torch.empty_like(a)
    _run("nsp_maxpool2d_relu_bwd", lib.nsp_maxpool2d_relu_bwd, int(a.dtype == torch.bfloat16), int(dy.dtype == torch.bfloat16),
         ptr(a), ptr(dy), ptr(dz), B, T, F, C, int(pool_t), int(pool_f), int(in_chmajor), current_stream_ptr())
    return dz


def conv3x3_wgrad(a, dz, dw, dbias, B, T, F, in_chmajor=False):
    """dw[CO,CI,3,3] += ..., dbias[CO] += ... (nsp_conv3x3_wgrad); a holds B*T*F*CI elements, dz `[B,T,F,CO]`."""
    _require_cuda(a, dz, dw)
    CO, CI = dw.shape[0], dw.shape[1]
    a, dz = a.contiguous(), dz.contiguous()
    assert a.numel() == B * T * F * CI and dz.numel() == B * T * F * CO and dw.is_contiguous()
    if CI == 32 and CO == 32 and not in_chmajor and a.dtype == torch.bfloat16 and dz.dtype == torch.bfloat16:
        _run("nsp_conv3x3_c32_wgrad_tc", lib.nsp_conv3x3_c32_wgrad_tc, ptr(a), ptr(dz), ptr(dw), B, T, F, current_stream_ptr(),
             flops=2.0 * B * T * F * 9 * CI * CO, tag="conv3x3_wgrad_tc")
        if dbias is not None:
            colsum_acc(dz.view(-1, CO), dbias)
        return dw
    _run("nsp_conv3x3_wgrad", lib.nsp_conv3x3_wgrad, int(a.dtype == torch.bfloat16), int(dz.dtype == torch.bfloat16), ptr(a),
         int(in_chmajor), ptr(dz), ptr(dw), ptr(dbias), B, T, F, CI, CO, current_stream_ptr(),
         flops=2.0 * B * T * F * 9 * CI * CO, tag="conv3x3_wgrad")
    return dw


def conv3x3_dgrad_weight(weight):
    """Taps of the input-gradient convolution: w'[ci, co, ky, kx] = w[co, ci, 2-ky, 2-kx]."""
    return weight.detach().flip(2, 3).transpose(0, 1).contiguous()


def relpos_attention_bwd(q, k, v, klens, n_heads, out, dout, r=None, u_bias=None, v_bias=None, clamp_len=-1, causal=False,
                         lookahead=0, chunk_c=0, chunk_l=0, dr=None, du=None, dvb=None, stats=None):
    """Backward of relpos_attention (nsp_relpos_attention_bwd).  Returns dqkv `[B, T, 3*D]` (requires Tq == Tk) with
    dq | dk | dv in the I/O dtype; dr `[rlen, D]`, du, dvb fp32 are accumulated in place when given."""
    _require_cuda(q, k, v, klens, out, dout)
    B, Tq, D = q.shape
    Tk = k.shape[1]
    assert Tq == Tk, "training path: no cache (Tq == Tk)"
    dk = D // n_heads
    dout = dout.to(q.dtype)
    dout = dout if dout.stride(2) == 1 and dout.stride(0) == Tq * dout.stride(1) else dout.contiguous()
    out = out if out.stride(2) == 1 and out.stride(0) == Tq * out.stride(1) else out.contiguous()
    if r is not None:
        r = r.reshape(-1, D) if r.dim() == 3 else r
    rlen = r.shape[0] if r is not None else 0
    if os.environ.get("NSP_DISABLE_ATTN_BWD_TC"):      # debugging switch: force the CUDA-core backward kernels
        stats = None
    dqkv = torch.empty(B, Tq, 3 * D, dtype=q.dtype, device=q.device)
    ws_bytes = lib.nsp_relpos_attention_bwd_workspace_bytes(B, n_heads, Tq, rlen, int(clamp_len), int(r is not None))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device)
    dq_, dk_, dv_ = dqkv[:, :, :D], dqkv[:, :, D:2 * D], dqkv[:, :, 2 * D:]
    _run("nsp_relpos_attention_bwd", lib.nsp_relpos_attention_bwd, int(q.dtype == torch.bfloat16), ptr(q), q.stride(1), ptr(k), k.stride(1),
         ptr(v), v.stride(1), ptr(r), r.stride(0) if r is not None else 0, rlen, ptr(u_bias), ptr(v_bias), ptr(klens),
         ptr(out), out.stride(1), ptr(dout), dout.stride(1), ptr(dq_), 3 * D, ptr(dk_), 3 * D, ptr(dv_), 3 * D,
         ptr(dr), dr.stride(0) if dr is not None else 0, ptr(du), ptr(dvb), ptr(stats), B, n_heads, Tq, Tk, dk, int(clamp_len),
         int(causal), int(lookahead), int(chunk_c), int(chunk_l), ptr(ws), ws_bytes, current_stream_ptr(),
         flops=16.0 * B * n_heads * Tq * Tk * dk, tag="attention_bwd")
    return dqkv


def conformer_conv_bwd(x, taps, dw_bias, norm_w, norm_b, eps, dy, dtaps, dbias, dnorm_w, dnorm_b, causal=False):
    """Backward of conformer_conv (LayerNorm variant; nsp_conformer_conv_bwd).  x, dy `[B,T,d]` in the I/O dtype;
    taps fp32 `[k,d]`; dtaps `[k,d]`, dbias, dnorm_w, dnorm_b fp32 `[d]` are accumulated.  Returns dx like x."""
    _require_cuda(x, dy)
    B, T, d = x.shape
    k = taps.shape[0]
    x = x if x.stride(2) == 1 and x.stride(0) == T * x.stride(1) else x.contiguous()
    dy = dy.to(x.dtype)
    dy = dy if dy.stride(2) == 1 and dy.stride(0) == T * dy.stride(1) else dy.contiguous()
    dz = torch.empty(B, T, d, dtype=x.dtype, device=x.device)
    dx = torch.empty(B, T, d, dtype=x.dtype, device=x.device)
    ws_bytes = lib.nsp_conformer_conv_bwd_workspace_bytes(B, T, d, k)
    ws = torch.empty(max(1, ws_bytes), dtype=torch.uint8, device=x.device)
    _run("nsp_conformer_conv_bwd", lib.nsp_conformer_conv_bwd, int(x.dtype == torch.bfloat16), ptr(x), x.stride(1), ptr(taps), ptr(dw_bias),
         0, ptr(norm_w), ptr(norm_b), float(eps), ptr(dy), dy.stride(1), ptr(dz), d, ptr(dx), d,
         ptr(dtaps), ptr(dbias), ptr(dnorm_w), ptr(dnorm_b), B, T, d, k, int(causal), ptr(ws), ws_bytes, current_stream_ptr())
    return dx
