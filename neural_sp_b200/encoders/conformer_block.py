"""Conformer encoder block (reference encoders/conformer_block.py:20-182), B200-native.

Macaron FFN -> rel-pos MHSA -> conv module -> FFN -> LayerNorm, pre-norm, fc_factor 0.5.  Parameter names
match the reference (norm1..5, feed_forward_macaron, self_attn, conv, feed_forward).  The fp32 residual
stream is updated in place by the GEMM epilogues; every LayerNorm writes the operand the next GEMM reads."""
import random

import torch
import torch.nn as nn

from .. import ops
from ..modules._prep import get_precision
from ..modules.conformer_convolution import ConformerConvBlock
from ..modules.positionwise_feed_forward import PositionwiseFeedForward as FFN
from ..modules.relative_multihead_attention import RelativeMultiheadAttentionMechanism as RelMHA

random.seed(1)


def _ln(norm, xs, prec):
    if prec == "bf16":
        return ops.layernorm(xs, norm.weight, norm.bias, norm.eps, out_fp32=False, out_bf16=True)
    return ops.layernorm(xs, norm.weight, norm.bias, norm.eps)


class ConformerEncoderBlock(nn.Module):
    def __init__(self, d_model, d_ff, n_heads, kernel_size, dropout, dropout_att, dropout_layer, layer_norm_eps,
                 ffn_activation, param_init, pe_type, clamp_len, ffn_bottleneck_dim, unidirectional,
                 normalization='layer_norm'):
        super().__init__()
        self.n_heads = n_heads
        self.fc_factor = 0.5
        self.norm1 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.feed_forward_macaron = FFN(d_model, d_ff, dropout, ffn_activation, param_init, ffn_bottleneck_dim)
        self.norm2 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.self_attn = RelMHA(kdim=d_model, qdim=d_model, adim=d_model, odim=d_model, n_heads=n_heads,
                                dropout=dropout_att, param_init=param_init, xl_like=pe_type == 'relative_xl',
                                clamp_len=clamp_len)
        self.norm3 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.conv = ConformerConvBlock(d_model, kernel_size, param_init, normalization, causal=unidirectional)
        self.conv_context = kernel_size
        self.norm4 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.feed_forward = FFN(d_model, d_ff, dropout, ffn_activation, param_init, ffn_bottleneck_dim)
        self.norm5 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.dropout = nn.Dropout(dropout)
        self.dropout_layer = dropout_layer
        self._xx_aws = None

    @property
    def xx_aws(self):
        return self._xx_aws

    def reset_visualization(self):
        self._xx_aws = None

    def forward(self, xs, klens, cache=None, pos_embs=None, rel_bias=(None, None), mask_kw=None):
        """xs fp32 `[B, T, d]` (updated in place), klens int32 `[B]` CUDA (valid keys INCLUDING cached frames),
        pos_embs fp32 `[>= n_cache + T, d]`.  cache (streaming, conformer_block.py:143-170): dict with the previous
        chunks' normalised attention input ``input_san`` `[B, n_cache, d]` and conv-module input ``input_conv``
        `[B, <= k + T - 1, d]`; the returned new_cache holds both extended by this chunk."""
        if self.training and (self.dropout.p > 0 or self.self_attn.dropout_attn.p > 0):
            raise NotImplementedError("dropout > 0 in train() mode runs through the autograd training path only (grad enabled); "
                                      "this is the inference-kernel path")
        prec = get_precision(self)
        mask_kw = mask_kw or {}
        u_bias, v_bias = rel_bias
        new_cache = {}
        qlen = xs.size(1)
        if self.dropout_layer > 0:          # LayerDrop (conformer_block.py:122-126): eval also rescales
            if self.training and random.random() < self.dropout_layer:
                return xs, new_cache
            ops.scale_(xs, 1.0 / (1 - self.dropout_layer))
        xs = self.feed_forward_macaron(_ln(self.norm1, xs, prec), residual=xs, scale=self.fc_factor, out=xs)
        h = _ln(self.norm2, xs, prec)
        kv = h if cache is None else torch.cat([cache['input_san'].to(h.dtype), h], dim=1)
        new_cache['input_san'] = kv
        kvc = cache.get('_kv') if cache is not None else None      # K / V projected in earlier chunks (ours, not the reference's)
        xs, new_cache['_kv'] = self.self_attn(kv, h, pos_embs, klens, u_bias, v_bias, residual=xs, out=xs, kv_cache=kvc,
                                              return_kv=True, **mask_kw)
        c_in = _ln(self.norm3, xs, prec)
        if cache is not None:               # left context of the depthwise convolution, restricted to the kernel size
            c_in = torch.cat([cache['input_conv'].to(c_in.dtype), c_in], dim=1)
            c_in = c_in[:, max(0, c_in.size(1) - (self.conv_context + qlen - 1)):].contiguous()
        new_cache['input_conv'] = c_in
        xs = self.conv(c_in, residual=xs, out=xs, keep_last=qlen if cache is not None else None)
        xs = self.feed_forward(_ln(self.norm4, xs, prec), residual=xs, scale=self.fc_factor, out=xs)
        xs = ops.layernorm(xs, self.norm5.weight, self.norm5.bias, self.norm5.eps)
        return xs, new_cache
