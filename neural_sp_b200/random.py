"""Random state of the library's dropout kernels (csrc/dropout.cu).

The state is two uint64 in DEVICE memory -- (seed, offset) -- read by every dropout kernel, plus a host-side counter
that hands every dropout call its own `stream` id (saved by the autograd node so that the backward regenerates the
forward's mask).  Eager training needs nothing else: ids never repeat.  A captured CUDA graph bakes its ids in, so a step
that is replayed must contain one `advance()` (a 1-thread kernel: offset += 1) to draw fresh masks on every replay."""
import torch

_state = {}          # device -> int64 tensor [2] = (seed, offset)
_seed = 0x5EED_B200
_next_stream = 0


def manual_seed(seed):
    """Re-seed (and rewind) the dropout generator on every device it has been used on."""
    global _seed, _next_stream
    _seed = int(seed) & 0x7FFF_FFFF_FFFF_FFFF
    _next_stream = 0
    for dev, t in _state.items():
        t.copy_(torch.tensor([_seed, 0], dtype=torch.int64), non_blocking=False)


def state(device):
    """int64 `[2]` tensor (seed, offset) on `device` (created on first use)."""
    device = torch.device(device)
    key = (device.type, device.index if device.index is not None else (torch.cuda.current_device() if device.type == "cuda" else 0))
    t = _state.get(key)
    if t is None:
        t = torch.tensor([_seed, 0], dtype=torch.int64).to(device)
        _state[key] = t
    return t


def next_stream():
    """A fresh call-site id (uint32)."""
    global _next_stream
    _next_stream = (_next_stream + 1) & 0xFFFF_FFFF
    return _next_stream


def advance(device):
    """offset += 1 on the device (enqueue once per training step inside a captured CUDA graph)."""
    from . import ops
    ops.rng_advance(state(device))
