"""Position-wise feed-forward layer (reference modules/positionwise_feed_forward.py:22-89), B200-native.

Same constructor and parameter names (``w_1``, ``w_2``).  ``forward`` here computes the whole residual
branch in two tcgen05 GEMMs with fused epilogues:  ``residual + scale * w_2(act(w_1(x)))``."""
import torch
import torch.nn as nn

from .. import ops
from ._prep import prepared, get_precision, act_dtype
from .initialization import init_with_xavier_uniform

_ACTS = {"relu": "relu", "swish": "swish", "gelu": "gelu", "gelu_accurate": "gelu_accurate", "glu": "glu"}


class LinearGLUBlock(nn.Module):
    """`F.glu(Linear(idim, 2*idim)(x))` (reference modules/glu.py:11-22); runs as one GEMM with a GLU epilogue."""

    def __init__(self, idim):
        super().__init__()
        self.fc = nn.Linear(idim, idim * 2)


class PositionwiseFeedForward(nn.Module):
    def __init__(self, d_model, d_ff, dropout, activation, param_init, bottleneck_dim=0):
        super().__init__()
        if bottleneck_dim > 0:
            raise NotImplementedError("low-rank FFN (ffn_bottleneck_dim > 0) is not on the B200 path yet")
        if activation not in _ACTS:
            raise NotImplementedError(activation)
        self.bottleneck_dim = 0
        self.act_name = _ACTS[activation]
        self.w_1 = nn.Linear(d_model, d_ff)
        self.w_2 = nn.Linear(d_ff, d_model)
        if activation == "glu":
            self.activation = LinearGLUBlock(d_ff)          # parameter name `activation.fc.*` as in the reference
        self.dropout = nn.Dropout(p=dropout)
        if param_init == 'xavier_uniform':
            for n, p in self.named_parameters():
                init_with_xavier_uniform(n, p)

    def forward(self, xs, residual=None, scale=1.0, out=None):
        """xs: normalised input (bf16 in bf16 mode, else fp32) `[B, T, d_model]`.
        Returns ``residual + scale * FFN(xs)`` (fp32), or ``scale * FFN(xs)`` when residual is None."""
        prec = get_precision(self)
        w1 = prepared(self, "w_1", prec, (self.w_1.weight,))
        w2 = prepared(self, "w_2", prec, (self.w_2.weight,))
        if self.act_name == "glu":
            h = ops.linear(xs, w1, self.w_1.bias, prec=prec, out_dtype=act_dtype(prec))
            wg = prepared(self, "glu_fc", prec, (self.activation.fc.weight,))
            h = ops.linear(h, wg, self.activation.fc.bias, prec=prec, glu=True, out_dtype=act_dtype(prec))
        else:
            h = ops.linear(xs, w1, self.w_1.bias, prec=prec, act=self.act_name, out_dtype=act_dtype(prec))
        return ops.linear(h, w2, self.w_2.bias, prec=prec, residual=residual, alpha=scale,
                          out_dtype=torch.float32, out=out)
