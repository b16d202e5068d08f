"""Conformer convolution module (reference modules/conformer_convolution.py:17-129), B200-native.

Same constructor and parameter names (``pointwise_conv1`` Conv1d(d,2d,1), ``depthwise_conv``
Conv1d(d,d,k,groups=d), ``norm``, ``pointwise_conv2``).  Computation stays time-major (no transposes):
pointwise_conv1+GLU is one GEMM with a GLU epilogue, depthwise conv + norm + Swish is one fused kernel,
pointwise_conv2 + residual is one GEMM."""
import torch
import torch.nn as nn

from .. import ops
from ._prep import prepared, cached, get_precision, act_dtype
from .initialization import init_with_xavier_uniform, init_with_lecun_normal


class ConformerConvBlock(nn.Module):
    def __init__(self, d_model, kernel_size, param_init, normalization='batch_norm', causal=False):
        super().__init__()
        assert (kernel_size - 1) % 2 == 0, 'kernel_size must be the odd number.'
        assert kernel_size >= 3, 'kernel_size must be larger than 3.'
        self.kernel_size = kernel_size
        self.causal = causal
        self.padding = (kernel_size - 1) if causal else (kernel_size - 1) // 2
        self.pointwise_conv1 = nn.Conv1d(d_model, d_model * 2, kernel_size=1, stride=1, padding=0)
        self.depthwise_conv = nn.Conv1d(d_model, d_model, kernel_size=kernel_size, stride=1, padding=self.padding,
                                        groups=d_model, bias=True)
        self.normalization = normalization
        if normalization == 'batch_norm':
            self.norm = nn.BatchNorm1d(d_model)
        elif normalization == 'group_norm':
            self.norm = nn.GroupNorm(num_groups=max(1, d_model // 2), num_channels=d_model)
        elif normalization == 'layer_norm':
            self.norm = nn.LayerNorm(d_model, eps=1e-12)
        else:
            raise NotImplementedError(normalization)
        self.pointwise_conv2 = nn.Conv1d(d_model, d_model, kernel_size=1, stride=1, padding=0)
        convs = (self.pointwise_conv1, self.pointwise_conv2, self.depthwise_conv)
        if param_init == 'xavier_uniform':
            for c in convs:
                for n, p in c.named_parameters():
                    init_with_xavier_uniform(n, p)
        elif param_init == 'lecun':
            for c in convs:
                for n, p in c.named_parameters():
                    init_with_lecun_normal(n, p, 0.1)

    def forward(self, xs, residual=None, out=None, keep_last=None):
        """xs `[B, T, d]` normalised input.  Returns ``residual + conv_module(xs)`` (fp32).
        keep_last = q (streaming, conformer_block.py:160-166): xs carries cached left context in front of the q new
        frames; only the last q frames go through pointwise_conv2 / the residual add."""
        prec = get_precision(self)
        if self.normalization == 'batch_norm' and self.training:
            raise NotImplementedError("BatchNorm statistics update (training mode) is not on the B200 path; "
                                      "use conformer_normalization=layer_norm as the LibriSpeech recipes do")
        if self.normalization == 'group_norm' and self.norm.num_groups * 2 != self.norm.num_channels:
            raise NotImplementedError("GroupNorm with other than 2 channels per group")
        w1 = prepared(self, "pw1", prec, (self.pointwise_conv1.weight,), build=lambda w: w.squeeze(-1))
        g = ops.linear(xs, w1, self.pointwise_conv1.bias, prec=prec, glu=True, out_dtype=act_dtype(prec))
        rm = getattr(self.norm, "running_mean", None)
        rv = getattr(self.norm, "running_var", None)
        taps = cached(self, "dw_taps", (self.depthwise_conv.weight,),
                      lambda w: w.reshape(w.size(0), -1).t().contiguous().float())          # [k, d]
        c = ops.conformer_conv(g, taps, self.depthwise_conv.bias, self.normalization,
                               self.norm.weight, self.norm.bias, self.norm.eps, rm, rv, causal=self.causal)
        if keep_last is not None and keep_last < c.size(1):
            c = c[:, c.size(1) - keep_last:].contiguous()
        w2 = prepared(self, "pw2", prec, (self.pointwise_conv2.weight,), build=lambda w: w.squeeze(-1))
        return ops.linear(c, w2, self.pointwise_conv2.bias, prec=prec, residual=residual, out_dtype=torch.float32, out=out)
