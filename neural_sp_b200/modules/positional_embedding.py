"""Positional tables (reference modules/positional_embedding.py): the TransformerXL-style sinusoid table used by
relative attention (built by a CUDA kernel, cached per length) and the absolute 'add' / 'none' encodings of the plain
Transformer encoder (the sinusoid buffer `pe` is part of the reference's state_dict, so it is kept as a buffer)."""
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


class PositionalEncoding(nn.Module):
    """Reference positional_embedding.py:20-95 for pe_type 'add' and 'none': ``xs * sqrt(d) [+ pe[offset:offset+T]]``.
    The '1dconv*' variants (a causal Conv1d stack) are not on the B200 path."""

    def __init__(self, d_model, dropout, pe_type, param_init, max_len=5000):
        super().__init__()
        if pe_type not in ('add', 'none'):
            raise NotImplementedError("positional encoding pe_type=%r is not on the B200 path "
                                      "(none, add, relative, relative_xl are)" % pe_type)
        self.d_model = d_model
        self.pe_type = pe_type
        self.scale = math.sqrt(d_model)
        if pe_type == 'add':
            pe = torch.zeros(max_len, d_model, dtype=torch.float32)
            position = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
            div_term = torch.exp(torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model))
            pe[:, 0::2] = torch.sin(position * div_term)
            pe[:, 1::2] = torch.cos(position * div_term)
            self.register_buffer('pe', pe.unsqueeze(0))
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, xs, scale=True, offset=0):
        """xs fp32 `[B, T, d]` contiguous, updated in place."""
        a = self.scale if scale else 1.0           # (the encoder applies self.dropout in its training path)
        if self.pe_type == 'none':
            return ops.scale_(xs, a) if a != 1.0 else xs
        T = xs.size(1)
        assert offset + T <= self.pe.size(1), "utterance longer than the positional table (max_len)"
        return ops.add_pos_enc_(xs, self.pe[0, offset:offset + T].contiguous(), a)
