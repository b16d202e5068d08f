"""Transformer encoder block (reference encoders/transformer_block.py:20-140), B200-native.

Pre-norm MHSA + FFN.  The reference's ``pe_type in ['relaive', 'relative_xl']`` typo (:46) is kept: only
``relative_xl`` yields relative attention in a Transformer block; ``relative`` silently uses plain MHA."""
import random

import torch
import torch.nn as nn

from .. import ops
from ..modules._prep import get_precision
from ..modules.multihead_attention import MultiheadAttentionMechanism as MHA
from ..modules.positionwise_feed_forward import PositionwiseFeedForward as FFN
from ..modules.relative_multihead_attention import RelativeMultiheadAttentionMechanism as RelMHA
from .conformer_block import _ln

random.seed(1)


class TransformerEncoderBlock(nn.Module):
    def __init__(self, d_model, d_ff, n_heads, dropout, dropout_att, dropout_layer, layer_norm_eps,
                 ffn_activation, param_init, pe_type, clamp_len, ffn_bottleneck_dim):
        super().__init__()
        self.n_heads = n_heads
        self.rel_attn = pe_type in ['relaive', 'relative_xl']
        self.norm1 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        mha = RelMHA if self.rel_attn else MHA
        self.self_attn = mha(kdim=d_model, qdim=d_model, adim=d_model, odim=d_model, n_heads=n_heads,
                             dropout=dropout_att, param_init=param_init, xl_like=pe_type == 'relative_xl',
                             clamp_len=clamp_len)
        self.norm2 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.feed_forward = FFN(d_model, d_ff, dropout, ffn_activation, param_init, ffn_bottleneck_dim)
        self.dropout = nn.Dropout(dropout)
        self.dropout_layer = dropout_layer
        self._xx_aws = None

    @property
    def xx_aws(self):
        return self._xx_aws

    def reset_visualization(self):
        self._xx_aws = None

    def forward(self, xs, klens, cache=None, pos_embs=None, rel_bias=(None, None), mask_kw=None):
        """cache (streaming, transformer_block.py:113-122): ``input_san`` `[B, n_cache, d]`, the previous chunks'
        normalised attention input; klens counts cached frames too."""
        if self.training and (self.dropout.p > 0 or self.self_attn.dropout_attn.p > 0):
            raise NotImplementedError("dropout > 0 in train() mode runs through the autograd training path only (grad enabled); "
                                      "this is the inference-kernel path")
        prec = get_precision(self)
        mask_kw = mask_kw or {}
        new_cache = {}
        if self.dropout_layer > 0:
            if self.training and random.random() < self.dropout_layer:
                return xs, new_cache
            ops.scale_(xs, 1.0 / (1 - self.dropout_layer))
        h = _ln(self.norm1, xs, prec)
        kv = h if cache is None else torch.cat([cache['input_san'].to(h.dtype), h], dim=1)
        new_cache['input_san'] = kv
        kvc = cache.get('_kv') if cache is not None else None      # K / V projected in earlier chunks (ours, not the reference's)
        if self.rel_attn:
            xs, new_cache['_kv'] = self.self_attn(kv, h, pos_embs, klens, rel_bias[0], rel_bias[1], residual=xs, out=xs,
                                                  kv_cache=kvc, return_kv=True, **mask_kw)
        else:
            xs, new_cache['_kv'] = self.self_attn(kv, h, klens, residual=xs, out=xs, kv_cache=kvc, return_kv=True, **mask_kw)
        xs = self.feed_forward(_ln(self.norm2, xs, prec), residual=xs, scale=1.0, out=xs)
        return xs, new_cache
