"""Relative-position multi-head self-attention (reference modules/relative_multihead_attention.py:21-220).

Same constructor / parameter names (``w_key, w_value, w_query, w_out[, w_pos]``, bias-free by default).
K, V, Q projections run as ONE tcgen05 GEMM over the concatenated weight, the score/softmax/PV part is
the flash-style kernel of the library, and ``w_out`` + residual is a GEMM epilogue.
Reference quirks kept on purpose: queries are projected from ``key[:, -qlen:]`` (:171); for
``xl_like=False`` positions are projected with ``w_value`` (:176); HeadDrop is not applied (:207-215)."""
import math

import torch
import torch.nn as nn

from .. import ops
from ._prep import prepared, get_precision, act_dtype


class RelativeMultiheadAttentionMechanism(nn.Module):
    def __init__(self, kdim, qdim, adim, odim, n_heads, dropout, dropout_head=0., bias=False, param_init='',
                 xl_like=False, clamp_len=-1):
        super().__init__()
        assert adim % n_heads == 0
        assert kdim == qdim
        self.d_k = adim // n_heads
        self.n_heads = n_heads
        self.scale = math.sqrt(self.d_k)
        self.xl_like = xl_like
        self.clamp_len = clamp_len
        self.dropout_attn = nn.Dropout(p=dropout)
        self.dropout_head = dropout_head
        self.w_key = nn.Linear(kdim, adim, bias=bias)
        self.w_value = nn.Linear(kdim, adim, bias=bias)
        self.w_query = nn.Linear(qdim, adim, bias=bias)
        self.w_out = nn.Linear(adim, odim, bias=bias)
        if xl_like:
            self.w_pos = nn.Linear(qdim, adim, bias=bias)
        if param_init == 'xavier_uniform':
            g = 1 / math.sqrt(2)
            for lin, gain in ((self.w_key, g), (self.w_value, g), (self.w_query, g), (self.w_out, 1.0)):
                nn.init.xavier_uniform_(lin.weight, gain=gain)
                if bias:
                    nn.init.constant_(lin.bias, 0.)
            if xl_like:
                nn.init.xavier_uniform_(self.w_pos.weight)
                if bias:
                    nn.init.constant_(self.w_pos.bias, 0.)

    def _qkv_bias(self):
        if self.w_key.bias is None:
            return None
        return torch.cat([self.w_query.bias, self.w_key.bias, self.w_value.bias]).detach()

    def forward(self, key, query, pos_embs, klens, u_bias=None, v_bias=None, residual=None, out=None,
                causal=False, lookahead=0, chunk_c=0, chunk_l=0, kv_cache=None, return_kv=False):
        """key `[B, mlen+qlen, d]` normalised input (bf16/fp32 by precision); query is its last qlen frames.
        pos_embs `[>=klen, d]` fp32 sinusoid table (row = distance).  klens int32 `[B]` on the GPU.
        Returns ``residual + w_out(attention)`` (fp32) -- the attention weights are never materialised.
        Streaming: kv_cache = (K, V) `[B, mlen, D]` projected in earlier chunks -- then only the qlen new frames go through
        the QKV GEMM (the reference re-projects the whole cache every chunk, :169-171; row-wise the result is identical);
        return_kv -> (out, (K, V)) with this chunk's rows appended."""
        prec = get_precision(self)
        B, klen, _ = key.shape
        qlen = query.shape[1]
        D = self.n_heads * self.d_k
        wqkv = prepared(self, "qkv", prec, (self.w_query.weight, self.w_key.weight, self.w_value.weight),
                        build=lambda q, k, v: torch.cat([q, k, v], dim=0))
        if kv_cache is not None:
            assert kv_cache[0].size(1) == klen - qlen
            qkv = ops.linear(query, wqkv, self._qkv_bias(), prec=prec, out_dtype=act_dtype(prec))   # `[B, qlen, 3D]`
            q = qkv[:, :, :D].contiguous()
            k = torch.cat([kv_cache[0], qkv[:, :, D:2 * D]], dim=1)
            v = torch.cat([kv_cache[1], qkv[:, :, 2 * D:]], dim=1)
        else:
            qkv = ops.linear(key, wqkv, self._qkv_bias(), prec=prec, out_dtype=act_dtype(prec))     # `[B, klen, 3D]`
            q = qkv[:, klen - qlen:, :D]
            if klen != qlen:
                q = q.contiguous()
            k, v = qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
        # projected position table; only distances 0..clamp_len are ever gathered when clamp_len > 0.
        # The sinusoid rows are constants, so R depends only on the projection weight: cached per weight version.
        nrows = min(klen, self.clamp_len + 1) if self.clamp_len > 0 else klen
        wp_lin = self.w_pos if self.xl_like else self.w_value
        ckey = (prec, nrows, wp_lin.weight.data_ptr(), wp_lin.weight._version, pos_embs.data_ptr())
        cached = self.__dict__.get("_r_cache")
        if cached is not None and cached[0] == ckey:
            r = cached[1]
        else:
            wp = prepared(self, "pos", prec, (wp_lin.weight,))
            r = ops.linear(pos_embs[:nrows], wp, wp_lin.bias, prec=prec, out_dtype=act_dtype(prec))   # `[nrows, D]`
            self.__dict__["_r_cache"] = (ckey, r)
        cv = ops.relpos_attention(q, k, v, klens, self.n_heads, r=r, u_bias=u_bias, v_bias=v_bias,
                                  clamp_len=self.clamp_len, causal=causal, lookahead=lookahead,
                                  chunk_c=chunk_c, chunk_l=chunk_l)
        wo = prepared(self, "w_out", prec, (self.w_out.weight,))
        y = ops.linear(cv, wo, self.w_out.bias, prec=prec, residual=residual, out_dtype=torch.float32, out=out)
        return (y, (k, v)) if return_kv else y
