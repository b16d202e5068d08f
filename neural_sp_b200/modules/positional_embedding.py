"""Positional tables (reference modules/positional_embedding.py).  Only the TransformerXL-style sinusoid
table used by relative attention is needed on this path; it is built by a CUDA kernel and cached per length."""
import math

import torch
import torch.nn as nn

from .. import ops


class XLPositionalEmbedding(nn.Module):
    """Reference positional_embedding.py:98-140: returns (xs * sqrt(d) if scale, table `[L, 1, d]`)."""

    def __init__(self, d_model, dropout):
        super().__init__()
        self.d_model = d_model
        self.scale = math.sqrt(d_model)
        inv_freq = 1 / (10000 ** (torch.arange(0.0, d_model, 2.0) / d_model))
        self.register_buffer("inv_freq", inv_freq)
        self.dropout = nn.Dropout(p=dropout)
        self._tab = None

    def table(self, rows):
        """fp32 `[rows, d]`, row r <-> position -(r+1) <-> relative distance r."""
        if self._tab is None or self._tab.shape[0] < rows or self._tab.device != self.inv_freq.device:
            self._tab = ops.xl_pos_table(self.inv_freq.float().contiguous(), max(rows, 64))
        return self._tab[:rows]

    def forward(self, xs, scale=False, n_cache=0):
        if scale:
            xs = ops.scale_(xs, self.scale) if (xs.dtype == torch.float32 and xs.is_contiguous()) else xs * self.scale
        return xs, self.table(xs.size(1) + n_cache).unsqueeze(1)
