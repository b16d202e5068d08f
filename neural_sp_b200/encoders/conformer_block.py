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
        """xs fp32 `[B, T, d]` (updated in place), klens int32 `[B]` CUDA, pos_embs fp32 `[>=T, d]`."""
        if cache is not None:
            raise NotImplementedError("streaming caches are a 'next' row (SURVEY.md 8f-4)")
        if self.training and (self.dropout.p > 0 or self.self_attn.dropout_attn.p > 0):
            raise NotImplementedError("dropout > 0 in training mode is not on the B200 path yet")
        prec = get_precision(self)
        mask_kw = mask_kw or {}
        u_bias, v_bias = rel_bias
        if self.dropout_layer > 0:          # LayerDrop (conformer_block.py:122-126): eval also rescales
            if self.training and random.random() < self.dropout_layer:
                return xs, {}
            ops.scale_(xs, 1.0 / (1 - self.dropout_layer))
        xs = self.feed_forward_macaron(_ln(self.norm1, xs, prec), residual=xs, scale=self.fc_factor, out=xs)
        h = _ln(self.norm2, xs, prec)
        xs = self.self_attn(h, h, pos_embs, klens, u_bias, v_bias, residual=xs, out=xs, **mask_kw)
        xs = self.conv(_ln(self.norm3, xs, prec), residual=xs, out=xs)
        xs = self.feed_forward(_ln(self.norm4, xs, prec), residual=xs, scale=self.fc_factor, out=xs)
        xs = ops.layernorm(xs, self.norm5.weight, self.norm5.bias, self.norm5.eps)
        return xs, {}
