"""Scaled-dot multi-head self-attention (reference modules/multihead_attention.py:17-157), encoder use only
(self-attention with key padding / causal / chunk masks).  Biases are on by default as in the reference."""
import math

import torch
import torch.nn as nn

from .. import ops
from ._prep import prepared, get_precision, act_dtype


class MultiheadAttentionMechanism(nn.Module):
    def __init__(self, kdim, qdim, adim, odim, n_heads, dropout, dropout_head=0., atype='scaled_dot', bias=True,
                 param_init='', xl_like=False, clamp_len=-1):
        super().__init__()
        if atype != 'scaled_dot':
            raise NotImplementedError("only scaled_dot attention is on the B200 encoder path")
        assert adim % n_heads == 0
        self.atype = atype
        self.d_k = adim // n_heads
        self.n_heads = n_heads
        self.scale = math.sqrt(self.d_k)
        self.dropout_attn = nn.Dropout(p=dropout)
        self.dropout_head = dropout_head
        self.w_key = nn.Linear(kdim, adim, bias=bias)
        self.w_value = nn.Linear(kdim, adim, bias=bias)
        self.w_query = nn.Linear(qdim, adim, bias=bias)
        self.w_out = nn.Linear(adim, odim, bias=bias)
        if param_init == 'xavier_uniform':
            g = 1 / math.sqrt(2)
            for lin, gain in ((self.w_key, g), (self.w_value, g), (self.w_query, g), (self.w_out, 1.0)):
                nn.init.xavier_uniform_(lin.weight, gain=gain)
                if bias:
                    nn.init.constant_(lin.bias, 0.)
        self.reset()

    def reset(self):
        self.key = None
        self.value = None
        self.mask = None

    def forward(self, key, query, klens, residual=None, out=None, causal=False, lookahead=0, chunk_c=0, chunk_l=0,
                kv_cache=None, return_kv=False):
        """Streaming: kv_cache = (K, V) projected in earlier chunks, return_kv -> (out, (K, V)); see RelMHA.forward."""
        prec = get_precision(self)
        B, klen, _ = key.shape
        qlen = query.shape[1]
        D = self.n_heads * self.d_k
        wqkv = prepared(self, "qkv", prec, (self.w_query.weight, self.w_key.weight, self.w_value.weight),
                        build=lambda q, k, v: torch.cat([q, k, v], dim=0))
        bias = None
        if self.w_key.bias is not None:
            bias = torch.cat([self.w_query.bias, self.w_key.bias, self.w_value.bias]).detach()
        if kv_cache is not None:
            assert kv_cache[0].size(1) == klen - qlen
            qkv = ops.linear(query, wqkv, bias, prec=prec, out_dtype=act_dtype(prec))
            q = qkv[:, :, :D].contiguous()
            k = torch.cat([kv_cache[0], qkv[:, :, D:2 * D]], dim=1)
            v = torch.cat([kv_cache[1], qkv[:, :, 2 * D:]], dim=1)
        else:
            qkv = ops.linear(key, wqkv, bias, prec=prec, out_dtype=act_dtype(prec))
            q = qkv[:, klen - qlen:, :D]
            if klen != qlen:
                q = q.contiguous()
            k, v = qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
        cv = ops.relpos_attention(q, k, v, klens, self.n_heads, r=None,
                                  causal=causal, lookahead=lookahead, chunk_c=chunk_c, chunk_l=chunk_l)
        wo = prepared(self, "w_out", prec, (self.w_out.weight,))
        y = ops.linear(cv, wo, self.w_out.bias, prec=prec, residual=residual, out_dtype=torch.float32, out=out)
        return (y, (k, v)) if return_kv else y
