"""Position-wise feed-forward layer (reference modules/positionwise_feed_forward.py:22-89), B200-native.

Same constructor and parameter names (``w_1``, ``w_2``; ``w_1_e / w_1_d / w_2_e / w_2_d`` for the low-rank form with
``bottleneck_dim > 0``, :41-45, :87).  ``forward`` here computes the whole residual branch in two (four) tcgen05 GEMMs with
fused epilogues:  ``residual + scale * w_2(act(w_1(x)))``."""
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
        if activation not in _ACTS:
            raise NotImplementedError(activation)
        self.bottleneck_dim = bottleneck_dim
        self.act_name = _ACTS[activation]
        if bottleneck_dim > 0:             # low-rank: every projection factored through `bottleneck_dim` (same order as the
            self.w_1_e = nn.Linear(d_model, bottleneck_dim)      # reference so that seeded init draws identically)
            self.w_1_d = nn.Linear(bottleneck_dim, d_ff)
            self.w_2_e = nn.Linear(d_ff, bottleneck_dim)
            self.w_2_d = nn.Linear(bottleneck_dim, d_model)
        else:
            self.w_1 = nn.Linear(d_model, d_ff)
            self.w_2 = nn.Linear(d_ff, d_model)
        if activation == "glu":
            self.activation = LinearGLUBlock(d_ff)          # parameter name `activation.fc.*` as in the reference
        self.dropout = nn.Dropout(p=dropout)
        if param_init == 'xavier_uniform':
            for n, p in self.named_parameters():
                init_with_xavier_uniform(n, p)

    @property
    def in_layers(self):
        """Linear layers from the input to the activation's input (the last one carries the activation epilogue)."""
        return [('w_1_e', self.w_1_e), ('w_1_d', self.w_1_d)] if self.bottleneck_dim > 0 else [('w_1', self.w_1)]

    @property
    def out_layers(self):
        """Linear layers from the activation's output back to d_model (the last one carries the residual epilogue)."""
        return [('w_2_e', self.w_2_e), ('w_2_d', self.w_2_d)] if self.bottleneck_dim > 0 else [('w_2', self.w_2)]

    @property
    def out_bias(self):
        return self.out_layers[-1][1].bias

    def forward(self, xs, residual=None, scale=1.0, out=None):
        """xs: normalised input (bf16 in bf16 mode, else fp32) `[B, T, d_model]`.
        Returns ``residual + scale * FFN(xs)`` (fp32), or ``scale * FFN(xs)`` when residual is None."""
        prec = get_precision(self)
        adt = act_dtype(prec)
        h = xs
        ins, outs = self.in_layers, self.out_layers
        for i, (name, lin) in enumerate(ins):
            act = self.act_name if (i == len(ins) - 1 and self.act_name != "glu") else None
            h = ops.linear(h, prepared(self, name, prec, (lin.weight,)), lin.bias, prec=prec, act=act, out_dtype=adt)
        if self.act_name == "glu":
            wg = prepared(self, "glu_fc", prec, (self.activation.fc.weight,))
            h = ops.linear(h, wg, self.activation.fc.bias, prec=prec, glu=True, out_dtype=adt)
        for name, lin in outs[:-1]:
            h = ops.linear(h, prepared(self, name, prec, (lin.weight,)), lin.bias, prec=prec, out_dtype=adt)
        name, lin = outs[-1]
        return ops.linear(h, prepared(self, name, prec, (lin.weight,)), lin.bias, prec=prec, residual=residual, alpha=scale,
                          out_dtype=torch.float32, out=out)
