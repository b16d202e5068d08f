"""Conformer encoder (reference encoders/conformer.py:18-111): TransformerEncoder with Conformer blocks
(version 1: FFN-MHSA(rel-pos)-Conv-FFN; ``conformer_v2`` encoder types: FFN-Conv-MHSA(plain)-FFN)."""
import copy

import torch.nn as nn

from .conformer_block import ConformerEncoderBlock
from .conformer_block_v2 import ConformerEncoderBlock_v2
from .transformer import TransformerEncoder


class ConformerEncoder(TransformerEncoder):
    def __init__(self, input_dim, enc_type, n_heads, kernel_size, normalization, n_layers, n_layers_sub1,
                 n_layers_sub2, d_model, d_ff, ffn_bottleneck_dim, ffn_activation, pe_type, layer_norm_eps,
                 last_proj_dim, dropout_in, dropout, dropout_att, dropout_layer, subsample, subsample_type,
                 n_stacks, n_splices, frontend_conv, task_specific_layer, param_init, clamp_len, lookahead,
                 chunk_size_left, chunk_size_current, chunk_size_right, streaming_type):
        super().__init__(input_dim, enc_type, n_heads, n_layers, n_layers_sub1, n_layers_sub2, d_model, d_ff,
                         ffn_bottleneck_dim, ffn_activation, pe_type, layer_norm_eps, last_proj_dim, dropout_in,
                         dropout, dropout_att, dropout_layer, subsample, subsample_type, n_stacks, n_splices,
                         frontend_conv, task_specific_layer, param_init, clamp_len, lookahead, chunk_size_left,
                         chunk_size_current, chunk_size_right, streaming_type)
        causal = self.unidir or (self.streaming_type == 'mask')
        if 'conformer_v2' in enc_type:          # conv module before a plain MHA (reference conformer.py:80-84)
            block = ConformerEncoderBlock_v2
        else:
            assert pe_type in ['relative', 'relative_xl']
            block = ConformerEncoderBlock
        self.layers = nn.ModuleList([copy.deepcopy(block(
            d_model, d_ff, n_heads, kernel_size, dropout, dropout_att, dropout_layer * (lth + 1) / n_layers,
            layer_norm_eps, ffn_activation, param_init, pe_type, clamp_len, ffn_bottleneck_dim, causal, normalization))
            for lth in range(n_layers)])
        for sub, nl in (('sub1', n_layers_sub1), ('sub2', n_layers_sub2)):      # reference :93-107
            if nl > 0 and task_specific_layer:
                setattr(self, 'layer_' + sub, block(
                    d_model, d_ff, n_heads, kernel_size, dropout, dropout_att, dropout_layer * nl / n_layers, layer_norm_eps,
                    ffn_activation, param_init, pe_type, clamp_len, ffn_bottleneck_dim, causal, normalization))
        self.reset_parameters(param_init)


# command-line contract of the reference (add_args / define_name static methods): see encoders/cli.py
from . import cli as _cli  # noqa: E402

ConformerEncoder.add_args = staticmethod(_cli.conformer_add_args)
ConformerEncoder.define_name = staticmethod(_cli.conformer_define_name)
