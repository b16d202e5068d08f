"""nn.Linear whose matmul runs on the library's tcgen05 GEMM (state_dict-compatible with nn.Linear).

Forward under autograd is supported through a custom Function whose backward also uses the same GEMM
(dX = dY W, dW = dY^T X), so heads such as the CTC output layers train through the CUDA path."""
import torch
import torch.nn as nn

from .. import ops
from ._prep import prepared, get_precision


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, module, prec):
        wp = prepared(module, "w", prec, (weight,))
        y = ops.linear(x, wp, bias, prec=prec, out_dtype=torch.float32)
        ctx.save_for_backward(x, weight)
        ctx.module, ctx.prec, ctx.has_bias = module, prec, bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        prec = ctx.prec
        gy2 = gy.reshape(-1, gy.shape[-1]).float().contiguous()
        x2 = x.reshape(-1, x.shape[-1]).float()
        gyo = ops.to_bf16(gy2) if prec == "bf16" else gy2      # one operand copy shared by the dgrad and wgrad GEMMs
        gx = gw = gb = None
        from .. import autograd as ag
        nb = weight.shape[0] if (ctx.has_bias and ctx.needs_input_grad[2]) else 0
        flat = torch.zeros(weight.numel() + nb, dtype=torch.float32, device=weight.device)   # this node's all-reduce bucket
        if ctx.needs_input_grad[0]:
            # dX[M,K] = dY[M,N] @ W[N,K]  ->  GEMM with "weight" W^T [K,N]
            wt = prepared(ctx.module, "wT", prec, (weight,), build=lambda w: w.t().contiguous())
            gx = ops.linear(gyo, wt, None, prec=prec, out_dtype=torch.float32).reshape(x.shape)
        if ctx.needs_input_grad[1]:
            # dW[N,K] = dY^T[N,M] @ X[M,K] on the MN-major wgrad kernel (no transposed copies)
            gw = flat[:weight.numel()].view(weight.shape)
            ops.linear_wgrad(gyo, x2, prec, gw)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = flat[weight.numel():]
            ops.colsum_acc(gy2, gb)
        if ag._GRAD_SYNC is not None:
            ag._GRAD_SYNC(flat)
        return gx, gw, gb, None, None


class Linear(nn.Linear):
    precision = "bf16"

    def forward(self, x):
        prec = get_precision(self)
        if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad):
            return _LinearFn.apply(x, self.weight, self.bias, self, prec)
        wp = prepared(self, "w", prec, (self.weight,))
        return ops.linear(x, wp, self.bias, prec=prec, out_dtype=torch.float32)
