"""Host -> device input staging (reference Speech2Text.encode speech2text.py:396-409: `np2tensor` per utterance = one pageable
H2D copy each, then `pad_list` = one more device kernel per utterance).  Here the list of `[T_b, F]` numpy feature matrices is
packed on the host into ONE pinned `[B, T_max, F]` buffer (reused across steps, grown on demand) and uploaded with ONE
asynchronous copy on the current stream; the lengths stay a CPU IntTensor, as every encoder's contract expects."""
import torch

_staging = {}       # (device, feature dim) -> [pinned host buffer, event of the last upload from it]


def pad_and_upload(xs, device, pad_value=0.0):
    """xs: list of float32 numpy arrays `[T_b, F]` -> (FloatTensor `[B, T_max, F]` on `device`, IntTensor `[B]` on CPU)."""
    B = len(xs)
    lens = [int(x.shape[0]) for x in xs]
    Tm, Fd = max(lens), int(xs[0].shape[1])
    device = torch.device(device)
    key = (str(device), Fd)
    slot = _staging.get(key)
    if slot is None or slot[0].numel() < B * Tm * Fd:
        buf = torch.empty(max(B * Tm * Fd, 1 << 20), dtype=torch.float32)
        slot = [buf.pin_memory() if device.type == "cuda" else buf, None]
        _staging[key] = slot
    if slot[1] is not None:
        slot[1].synchronize()           # the previous upload has left the staging buffer before it is overwritten
    host = slot[0][:B * Tm * Fd].view(B, Tm, Fd)
    hn = host.numpy()
    for b, x in enumerate(xs):
        hn[b, :lens[b]] = x
        if lens[b] < Tm:
            hn[b, lens[b]:] = pad_value
    dev = torch.empty(B, Tm, Fd, dtype=torch.float32, device=device)
    dev.copy_(host, non_blocking=True)
    if device.type == "cuda":
        slot[1] = torch.cuda.Event()
        slot[1].record()
    return dev, torch.IntTensor(lens)
