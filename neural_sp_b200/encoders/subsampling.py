"""Intermediate time subsamplers (reference encoders/subsampling.py:13-246), B200-native: pooling variants run on
one CUDA kernel, the two parametric variants (concat / conv1d) run as tcgen05 GEMMs with a fused ReLU."""
import math

import torch
import torch.nn as nn

from .. import ops
from ..modules._prep import prepared, get_precision


def update_lens_pool1d(xlens, factor):
    """ceil-mode pool, kernel=stride=factor: n -> (n + 1 - factor) // factor + 1  (conv.py:443-445)."""
    return torch.IntTensor([(int(n) + 1 - factor) // factor + 1 for n in xlens])


class _PoolSubsampler(nn.Module):
    mode = "max"

    def __init__(self, subsampling_factor):
        super().__init__()
        self.factor = subsampling_factor

    def _lens(self, xlens):
        return update_lens_pool1d(xlens, self.factor)

    def forward(self, xs, xlens, batch_first=True):
        if self.factor == 1:
            return xs, xlens
        assert batch_first
        return ops.pool_time(xs, self.factor, self.mode), self._lens(xlens)


class MaxPoolSubsampler(_PoolSubsampler):       # reference :175-209
    mode = "max"


class MeanPoolSubsampler(_PoolSubsampler):      # reference :212-246
    mode = "mean"

    def _lens(self, xlens):
        # the reference's update_lens_1d takes the FLOOR formula for nn.AvgPool1d (conv.py:446-450: only MaxPool1d gets
        # the ceil-mode branch) although the pooled tensor itself has ceil(T / f) frames
        return torch.IntTensor([(int(n) - self.factor) // self.factor + 1 for n in xlens])


class DropSubsampler(_PoolSubsampler):          # reference :97-126
    mode = "drop"

    def _lens(self, xlens):
        return torch.IntTensor([max(1, math.ceil(int(n) / self.factor)) for n in xlens])


class AddSubsampler(_PoolSubsampler):           # reference :129-172
    mode = "add"

    def __init__(self, subsampling_factor):
        super().__init__(subsampling_factor)
        assert subsampling_factor <= 2

    def _lens(self, xlens):
        return torch.IntTensor([max(1, math.ceil(int(n) / self.factor)) for n in xlens])


class ConcatSubsampler(nn.Module):
    """Concatenate `factor` successive frames, project back with ReLU (reference :13-52); trailing frames are dropped."""

    def __init__(self, subsampling_factor, n_units):
        super().__init__()
        self.factor = subsampling_factor
        if subsampling_factor > 1:
            self.proj = nn.Linear(n_units * subsampling_factor, n_units)

    def forward(self, xs, xlens, batch_first=True):
        if self.factor == 1:
            return xs, xlens
        assert batch_first
        prec = get_precision(self)
        B, T, D = xs.shape
        To = T // self.factor
        x = xs[:, :To * self.factor].reshape(B, To, self.factor * D)
        y = ops.linear(x, prepared(self, "proj", prec, (self.proj.weight,)), self.proj.bias, prec=prec, act="relu",
                       out_dtype=torch.float32)
        return y, self._lens(xlens)

    def _lens(self, xlens):
        return torch.IntTensor([max(1, int(n) // self.factor) for n in xlens])


class Conv1dSubsampler(nn.Module):
    """Strided 'same' Conv1d + ReLU (reference :55-94) as a GEMM over gathered [t-1, t, t+1] frames."""

    def __init__(self, subsampling_factor, n_units, kernel_size=3):
        super().__init__()
        assert kernel_size % 2 == 1
        self.factor = subsampling_factor
        self.kernel_size = kernel_size
        if subsampling_factor > 1:
            self.conv1d = nn.Conv1d(n_units, n_units, kernel_size, stride=subsampling_factor, padding=(kernel_size - 1) // 2)

    def forward(self, xs, xlens, batch_first=True):
        if self.factor == 1:
            return xs, xlens
        assert batch_first
        prec = get_precision(self)
        B, T, D = xs.shape
        k, pad, f = self.kernel_size, (self.kernel_size - 1) // 2, self.factor
        To = (T + 2 * pad - (k - 1) - 1) // f + 1
        xp = torch.nn.functional.pad(xs, (0, 0, pad, pad))                       # layout plumbing only
        cols = xp.unfold(1, k, f)[:, :To].permute(0, 1, 3, 2).reshape(B, To, k * D)   # [B, To, k*D], index j*D + c
        w = prepared(self, "conv1d", prec, (self.conv1d.weight,), build=lambda w_: w_.permute(0, 2, 1).reshape(w_.size(0), -1))
        y = ops.linear(cols, w, self.conv1d.bias, prec=prec, act="relu", out_dtype=torch.float32)
        return y, self._lens(xlens)

    def _lens(self, xlens):
        k, pad, f = self.kernel_size, (self.kernel_size - 1) // 2, self.factor
        return torch.IntTensor([(int(n) + 2 * pad - (k - 1) - 1) // f + 1 for n in xlens])
