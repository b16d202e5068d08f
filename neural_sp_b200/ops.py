"""Thin tensor-level wrappers over the C ABI (include/nsp_b200.h).

Each function takes CUDA tensors, allocates outputs/workspaces through torch's caching allocator on
the current stream, and enqueues the CUDA kernels through ctypes.  No arithmetic happens in Python.
"""
import torch

from . import _lib
from ._lib import lib, check, ptr, current_stream_ptr


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.NspError("neural_sp_b200 ops need CUDA tensors (no CPU fallback); got device %s" % t.device)


def pack_labels(ys, device):
    """List of label lists -> (labels int32 [B, Lmax] on device, ylens int32 [B] on device, Lmax)."""
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
    check(lib.nsp_ctc_loss_fwd_bwd(ptr(logits), logits.stride(0), logits.stride(1), B, T, V,
                                   ptr(labels), Lmax, ptr(elens), ptr(ylens), int(blank), float(lsm_prob),
                                   ptr(nll), ptr(loss), ptr(grad), ptr(ws), ws_bytes, current_stream_ptr()),
          "nsp_ctc_loss_fwd_bwd")
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
    check(lib.nsp_ctc_forced_align(ptr(logits), B, T, V, ptr(labels), Lmax, ptr(elens), ptr(ylens), int(blank),
                                   ptr(trig), ptr(ws), ws_bytes, current_stream_ptr()),
          "nsp_ctc_forced_align")
    return trig
