"""Training path of the encoder: hand-written backward chains wired into torch.autograd.

The reference trains through torch autograd over its nn.Modules (encoders/conformer_block.py:95-182,
encoders/transformer_block.py, encoders/conv.py:347-396, modules/*).  Here every encoder block, the CNN
front-end, the time max-pool and the final LayerNorm is ONE ``torch.autograd.Function`` whose forward runs
the same fused CUDA kernels as inference (out of place, keeping what the backward needs) and whose backward
is a hand-written chain of CUDA kernels (tcgen05 dgrad / wgrad GEMMs, flash attention backward, fused
LayerNorm / depthwise-conv backward).  PyTorch only orders the Functions and accumulates ``.grad``.

Python here is orchestration only: no arithmetic on tensors outside the library kernels, except views,
allocation and the handful of `torch.zeros` gradient buffers.
"""
import torch

from . import ops
from . import random as nrandom
from .modules._prep import prepared, cached, act_dtype


def training_enabled(module):
    """The autograd path is taken iff grad mode is on and the module is in train() mode."""
    return torch.is_grad_enabled() and module.training


class _Grads:
    """fp32 gradient buffers of one autograd node, carved out of ONE zero-initialised flat tensor (one memset per node
    instead of one per parameter).  The flat tensor is also the node's all-reduce bucket: when a gradient-sync hook is
    installed (``set_grad_sync``), the node hands it over as soon as its backward has been enqueued, so the NCCL
    all-reduce of block k overlaps the backward of blocks k-1, k-2, ... (the returned ``.grad`` tensors are views of it)."""

    def __init__(self, params=()):
        self.g = {}
        self.flat = None
        params = [p for p in params if p is not None]
        self.params = params
        if params:
            total = sum(p.numel() for p in params)
            self.flat = torch.zeros(total, dtype=torch.float32, device=params[0].device)
            off = 0
            for p in params:
                if id(p) not in self.g:
                    self.g[id(p)] = self.flat[off:off + p.numel()].view(p.shape)
                    off += p.numel()

    def buf(self, p):
        t = self.g.get(id(p))
        if t is None:
            t = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
            self.g[id(p)] = t
        return t

    def get(self, p):
        return self.g.get(id(p))

    def put(self, p, value):
        """Store a gradient computed in another layout (tiny tensors only: one strided copy)."""
        self.buf(p).copy_(value.reshape(p.shape))

    def fused(self, params):
        """One [sum rows, cols] view over parameters that are adjacent in the flat buffer (fused QKV weight gradient)."""
        views = [self.g[id(p)] for p in params]
        rows = sum(v.shape[0] for v in views)
        first = views[0]
        ok = self.flat is not None and all(
            views[i + 1].data_ptr() == views[i].data_ptr() + views[i].numel() * 4 for i in range(len(views) - 1))
        if not ok:
            return None
        off = (first.data_ptr() - self.flat.data_ptr()) // 4
        return self.flat[off:off + rows * first[0].numel()].view(rows, *first.shape[1:])

    def done(self):
        if _GRAD_SYNC is not None and self.flat is not None:
            ops.wgrad_join()                       # the bucket must be complete before it is handed to the all-reduce
            # The hook may start an ASYNCHRONOUS in-place all-reduce of the bucket the returned .grad tensors are views of.
            # If a parameter already holds a gradient (gradient accumulation, or a parameter shared by several nodes such as
            # the relative_xl u_bias / v_bias), autograd's AccumulateGrad would read the bucket while NCCL is reducing it.
            for p in self.params:
                if p.grad is not None:
                    raise RuntimeError("neural_sp_b200: a gradient-sync hook is installed but a parameter of this node already "
                                       "has .grad (accumulation or a parameter shared between nodes): set grads to None before "
                                       "every backward, or reduce after the backward pass (dist.flat_allreduce_grads)")
            _GRAD_SYNC(self.flat)


_GRAD_SYNC = None


def set_grad_sync(fn):
    """Install (or clear with None) the hook called with every node's flat gradient bucket right after its backward."""
    global _GRAD_SYNC
    _GRAD_SYNC = fn


def _wT(module, name, prec, params, build=None):
    """Prepared TRANSPOSED weight (operand of the input-gradient GEMM dx = dy W)."""
    if build is None:
        return prepared(module, name + ".T", prec, params, build=lambda w: w.reshape(w.shape[0], -1).t().contiguous())
    return prepared(module, name + ".T", prec, params, build=lambda *ws: build(*ws).t().contiguous())


def _as2d(w):
    return w.reshape(w.shape[0], -1)


def _gop(x, prec):
    """Gradient tensor in the GEMM operand dtype (bf16 copy in bf16 mode)."""
    return ops.to_bf16(x.contiguous()) if prec == "bf16" else x


# ------------------------------------------------------------------------------------------------
# residual branches  y = x + scale * F(LN(x))
# ------------------------------------------------------------------------------------------------
def _ln_fwd(norm, x, prec):
    if prec == "bf16":
        return ops.layernorm(x, norm.weight, norm.bias, norm.eps, out_fp32=False, out_bf16=True)
    return ops.layernorm(x, norm.weight, norm.bias, norm.eps)


def _ln_bwd(norm, dn, x, dres, G, prec, next_bias=None, next_alpha=1.0):
    """-> (dx fp32, dx in the GEMM operand dtype): in bf16 mode the kernel also emits the bf16 copy the next branch's
    dgrad / wgrad GEMMs read, so no separate cast pass is needed.  next_bias: bias parameter of the residual-branch output
    layer that will receive dx as its dy (its gradient = alpha * column sums of dx, accumulated by the same kernel)."""
    dcol = G.buf(next_bias) if next_bias is not None else None
    if prec == "bf16":
        return ops.layernorm_bwd(dn, x, norm.weight, norm.eps, dres=dres, dgamma=G.buf(norm.weight), dbeta=G.buf(norm.bias),
                                 want_bf16=True, dcol=dcol, dcol_alpha=next_alpha)
    dx = ops.layernorm_bwd(dn, x, norm.weight, norm.eps, dres=dres, dgamma=G.buf(norm.weight), dbeta=G.buf(norm.bias),
                           dcol=dcol, dcol_alpha=next_alpha)
    return dx, dx


def _fusable(t):
    return t.dtype == torch.bfloat16 and t.shape[-1] % 16 == 0


def _drop_out(t_fn, x, p_res, alpha):
    """Residual-branch output  y = x + alpha * dropout(t)  (conformer_block.py:133-134 etc.): t_fn(fused) produces either the
    GEMM with the residual add fused in its epilogue (p_res == 0) or the bare branch output for the dropout-add kernel.
    Returns (y, stream id of the mask or 0)."""
    if p_res <= 0:
        return t_fn(True), 0
    sid = nrandom.next_stream()
    return ops.dropout_add(t_fn(False), x, p_res, alpha, sid), sid


def _drop_grad(dy, dyo, p_res, sid, prec):
    """Gradient entering a residual branch whose output was dropped: regenerate the mask on dy.
    -> (gradient for the bias column sums, gradient in the GEMM operand dtype, bias-fusion still valid?)"""
    if p_res <= 0:
        return dy, dyo, True
    dt = ops.dropout(dy, p_res, sid, out_dtype=act_dtype(prec))
    return dt, dt, False


def ffn_fwd(ffn, norm, x, scale, prec, p_res=0.0):
    """y = x + scale * dropout(out(dropout(act(in(LN(x))))));  in / out = one Linear each, or two for the low-rank form
    (positionwise_feed_forward.py:41-45, :87); GLU = an extra Linear with the gate fused in its epilogue (modules/glu.py)."""
    n = _ln_fwd(norm, x, prec)
    adt = act_dtype(prec)
    ins, outs = ffn.in_layers, ffn.out_layers
    glu = ffn.act_name == "glu"
    xin, h, z = [], n, None                      # xin[i] = input of ins[i] (operand of its weight gradient)
    for i, (name, lin) in enumerate(ins):
        xin.append(h)
        w = prepared(ffn, name, prec, (lin.weight,))
        if i == len(ins) - 1 and not glu:
            h, z = ops.linear(h, w, lin.bias, prec=prec, act=ffn.act_name, out_dtype=adt, save_pre=True)
        else:
            h = ops.linear(h, w, lin.bias, prec=prec, out_dtype=adt)
    h1 = None
    if glu:
        h1 = h
        wg = prepared(ffn, "glu_fc", prec, (ffn.activation.fc.weight,))
        h, z = ops.linear(h1, wg, ffn.activation.fc.bias, prec=prec, glu=True, out_dtype=adt, save_pre=True)
    p_h, sid_h = ffn.dropout.p, 0
    if p_h > 0:                    # dropout(act(.))  (positionwise_feed_forward.py:83)
        sid_h = nrandom.next_stream()
        h = ops.dropout(h, p_h, sid_h, inplace=True)
    xout = []                                    # xout[i] = input of outs[i]
    for name, lin in outs[:-1]:
        xout.append(h)
        h = ops.linear(h, prepared(ffn, name, prec, (lin.weight,)), lin.bias, prec=prec, out_dtype=adt)
    xout.append(h)
    name, lin = outs[-1]
    wl = prepared(ffn, name, prec, (lin.weight,))
    y, sid_r = _drop_out(lambda fused: ops.linear(h, wl, lin.bias, prec=prec, residual=x if fused else None,
                                                  alpha=scale if fused else 1.0,
                                                  out_dtype=torch.float32 if fused else adt), x, p_res, scale)
    return y, (x, xin, z, h1, xout, (p_h, sid_h, p_res, sid_r))


def ffn_bwd(ffn, norm, saved, dy, dyo, scale, prec, G, bias_done=False, nxt=(None, 1.0)):
    """bias_done: the producer of dy already accumulated the last layer's bias gradient; nxt = (bias param, alpha) of the
    branch that consumes this function's dx."""
    x, xin, z, h1, xout, (p_h, sid_h, p_res, sid_r) = saved
    adt = act_dtype(prec)
    ins, outs = ffn.in_layers, ffn.out_layers
    dyb, d, fused_ok = _drop_grad(dy, dyo, p_res, sid_r, prec)
    # output side, last layer first; `alpha` carries the branch scale through the first (= last forward) layer only
    alpha = scale
    for i in range(len(outs) - 1, -1, -1):
        name, lin = outs[i]
        ops.linear_wgrad(d, xout[i], prec, G.buf(lin.weight), alpha=alpha, side=True)
        if i == len(outs) - 1:
            if not (bias_done and fused_ok):
                ops.colsum_acc(dyb, G.buf(lin.bias), alpha=alpha)
        else:
            ops.colsum_acc(d, G.buf(lin.bias))
        d = ops.linear(d, _wT(ffn, name, prec, (lin.weight,)), None, prec=prec, alpha=alpha, out_dtype=adt)
        alpha = 1.0
    dh = d
    if p_h > 0:
        dh = ops.dropout(dh, p_h, sid_h, inplace=True)
    last_in = ins[-1][1]
    if ffn.act_name == "glu":
        fc = ffn.activation.fc
        if _fusable(z) and z.shape[-1] % 32 == 0:
            dpre = ops.act_bwd_bias(dh, z, None, G.buf(fc.bias), glu=True)
        else:
            dpre = ops.glu_bwd(dh, z)
            ops.colsum_acc(dpre, G.buf(fc.bias))
        ops.linear_wgrad(dpre, h1, prec, G.buf(fc.weight))
        dz = ops.linear(dpre, _wT(ffn, "glu_fc", prec, (fc.weight,)), None, prec=prec, out_dtype=adt)
        ops.colsum_acc(dz, G.buf(last_in.bias))
    elif _fusable(z):
        dz = ops.act_bwd_bias(dh, z, ffn.act_name, G.buf(last_in.bias))
    else:
        dz = ops.act_bwd(dh, z, ffn.act_name)
        ops.colsum_acc(dz, G.buf(last_in.bias))
    d = dz
    for i in range(len(ins) - 1, -1, -1):
        name, lin = ins[i]
        if i < len(ins) - 1:
            ops.colsum_acc(d, G.buf(lin.bias))
        ops.linear_wgrad(d, xin[i], prec, G.buf(lin.weight), side=True)
        d = ops.linear(d, _wT(ffn, name, prec, (lin.weight,)), None, prec=prec, out_dtype=torch.float32 if i == 0 else adt)
    return _ln_bwd(norm, d, x, dy, G, prec, nxt[0], nxt[1])


def _qkv_weight(attn, prec, transposed=False):
    params = (attn.w_query.weight, attn.w_key.weight, attn.w_value.weight)
    cat = lambda q, k, v: torch.cat([q, k, v], dim=0)     # noqa: E731
    if transposed:
        return _wT(attn, "qkv", prec, params, build=cat)
    return prepared(attn, "qkv", prec, params, build=cat)


def attn_fwd(attn, norm, x, pos, klens, u_bias, v_bias, mask_kw, prec, rel, p_res=0.0):
    n = _ln_fwd(norm, x, prec)
    D = attn.n_heads * attn.d_k
    bias = None
    if attn.w_key.bias is not None:
        bias = torch.cat([attn.w_query.bias, attn.w_key.bias, attn.w_value.bias]).detach()
    qkv = ops.linear(n, _qkv_weight(attn, prec), bias, prec=prec, out_dtype=act_dtype(prec))
    r, nrows = None, 0
    if rel:
        T = x.shape[1]
        nrows = min(T, attn.clamp_len + 1) if attn.clamp_len > 0 else T
        wp_lin = attn.w_pos if attn.xl_like else attn.w_value
        r = ops.linear(pos[:nrows], prepared(attn, "pos", prec, (wp_lin.weight,)), wp_lin.bias, prec=prec,
                       out_dtype=act_dtype(prec))
    clamp = attn.clamp_len if rel else -1
    cv, stats = ops.relpos_attention(qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:], klens, attn.n_heads, r=r,
                                     u_bias=u_bias if rel else None, v_bias=v_bias if rel else None, clamp_len=clamp,
                                     want_stats=True, **mask_kw)
    wo = prepared(attn, "w_out", prec, (attn.w_out.weight,))
    y, sid_r = _drop_out(lambda fused: ops.linear(cv, wo, attn.w_out.bias, prec=prec, residual=x if fused else None,
                                                  out_dtype=torch.float32 if fused else act_dtype(prec)), x, p_res, 1.0)
    return y, (x, n, qkv, r, cv, nrows, stats, (p_res, sid_r))


def attn_bwd(attn, norm, saved, dy, dyo, pos, klens, u_bias, v_bias, mask_kw, prec, rel, G, enc_bias_params,
             bias_done=False, nxt=(None, 1.0)):
    x, n, qkv, r, cv, nrows, stats, (p_res, sid_r) = saved
    D = attn.n_heads * attn.d_k
    d_in = x.shape[-1]
    dyb, dyo, fused_ok = _drop_grad(dy, dyo, p_res, sid_r, prec)
    ops.linear_wgrad(dyo, cv, prec, G.buf(attn.w_out.weight), side=True)
    if attn.w_out.bias is not None and not (bias_done and fused_ok):
        ops.colsum_acc(dyb, G.buf(attn.w_out.bias))
    dcv = ops.linear(dyo, _wT(attn, "w_out", prec, (attn.w_out.weight,)), None, prec=prec, out_dtype=act_dtype(prec))
    dr = torch.zeros(nrows, D, dtype=torch.float32, device=x.device) if rel else None
    du = dvb = None
    if rel and u_bias is not None:
        du, dvb = G.buf(enc_bias_params[0]).view(-1), G.buf(enc_bias_params[1]).view(-1)
    clamp = attn.clamp_len if rel else -1
    dqkv = ops.relpos_attention_bwd(qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:], klens, attn.n_heads, cv, dcv, r=r,
                                    u_bias=u_bias if rel else None, v_bias=v_bias if rel else None, clamp_len=clamp,
                                    dr=dr, du=du, dvb=dvb, stats=stats, **mask_kw)
    qkv_lins = (attn.w_query, attn.w_key, attn.w_value)
    gw = G.fused([lin.weight for lin in qkv_lins])             # the three weights are adjacent in the node's flat buffer
    if gw is not None:
        ops.linear_wgrad(dqkv, n, prec, gw, side=True)
    else:
        gw = torch.zeros(3 * D, d_in, dtype=torch.float32, device=x.device)
        ops.linear_wgrad(dqkv, n, prec, gw)
        for i, lin in enumerate(qkv_lins):
            G.put(lin.weight, gw[i * D:(i + 1) * D])
    if attn.w_key.bias is not None:
        gb = torch.zeros(3 * D, dtype=torch.float32, device=x.device)
        ops.colsum_acc(dqkv, gb)
        for i, lin in enumerate(qkv_lins):
            G.put(lin.bias, gb[i * D:(i + 1) * D])
    if rel:   # r = pos[:nrows] @ Wp^T (+ bp): the position projection shares w_value when not xl_like (reference :176)
        wp_lin = attn.w_pos if attn.xl_like else attn.w_value
        ops.linear_wgrad(dr, pos[:nrows], prec, G.buf(wp_lin.weight))
        if wp_lin.bias is not None:
            ops.colsum_acc(dr, G.buf(wp_lin.bias))
    dn = ops.linear(dqkv, _qkv_weight(attn, prec, transposed=True), None, prec=prec, out_dtype=torch.float32)
    return _ln_bwd(norm, dn, x, dy, G, prec, nxt[0], nxt[1])


def _bn_batch_stats(bn, stats, M):
    """(mean, biased var) fp32 `[d]` of the batch from (sum, sum of squares), and nn.BatchNorm1d's training-mode update of the
    running statistics (momentum, UNBIASED variance, num_batches_tracked); d-length vector arithmetic, all on the device."""
    if bn.momentum is None:
        raise NotImplementedError("BatchNorm with momentum=None (cumulative average) is not on the B200 path")
    s = stats.double()
    mean = s[0] / M
    var = (s[1] / M - mean * mean).clamp_min(0.0)
    with torch.no_grad():
        if bn.track_running_stats and bn.running_mean is not None:
            m = bn.momentum
            bn.running_mean.mul_(1 - m).add_((m * mean).to(bn.running_mean.dtype))
            bn.running_var.mul_(1 - m).add_((m * var * (M / max(M - 1, 1))).to(bn.running_var.dtype))
            bn.num_batches_tracked += 1
    return mean.float().contiguous(), var.float().contiguous()


def convmod_fwd(conv, norm, x, prec, p_res=0.0):
    if conv.normalization == 'group_norm' and conv.norm.num_groups * 2 != conv.norm.num_channels:
        raise NotImplementedError("GroupNorm with other than 2 channels per group")
    n = _ln_fwd(norm, x, prec)
    w1 = prepared(conv, "pw1", prec, (conv.pointwise_conv1.weight,), build=lambda w: w.squeeze(-1))
    g, pre = ops.linear(n, w1, conv.pointwise_conv1.bias, prec=prec, glu=True, out_dtype=act_dtype(prec), save_pre=True)
    taps = cached(conv, "dw_taps", (conv.depthwise_conv.weight,), lambda w: w.reshape(w.size(0), -1).t().contiguous().float())
    bn = None
    if conv.normalization == 'batch_norm':
        # statistics over ALL B*T frames, padded ones included (conformer_convolution.py:118-124); the normalised output comes
        # from the same fused kernel as inference, fed with the batch statistics
        z, stats = ops.dwconv_stats(g, taps, conv.depthwise_conv.bias, causal=conv.causal)
        mean, var = _bn_batch_stats(conv.norm, stats, g.shape[0] * g.shape[1])
        c = ops.conformer_conv(g, taps, conv.depthwise_conv.bias, 'batch_norm', conv.norm.weight, conv.norm.bias, conv.norm.eps,
                               mean, var, causal=conv.causal)
        bn = (z, mean, var)
    elif conv.normalization == 'group_norm':
        # per-frame pair statistics: the forward is the inference kernel as is; the backward wants z (no batch coupling)
        z, _ = ops.dwconv_stats(g, taps, conv.depthwise_conv.bias, causal=conv.causal)
        c = ops.conformer_conv(g, taps, conv.depthwise_conv.bias, 'group_norm', conv.norm.weight, conv.norm.bias, conv.norm.eps,
                               None, None, causal=conv.causal)
        bn = (z, None, None)
    else:
        c = ops.conformer_conv(g, taps, conv.depthwise_conv.bias, 'layer_norm', conv.norm.weight, conv.norm.bias, conv.norm.eps,
                               None, None, causal=conv.causal)
    w2 = prepared(conv, "pw2", prec, (conv.pointwise_conv2.weight,), build=lambda w: w.squeeze(-1))
    y, sid_r = _drop_out(lambda fused: ops.linear(c, w2, conv.pointwise_conv2.bias, prec=prec, residual=x if fused else None,
                                                  out_dtype=torch.float32 if fused else act_dtype(prec)), x, p_res, 1.0)
    return y, (x, n, pre, g, c, taps, bn, (p_res, sid_r))


def convmod_bwd(conv, norm, saved, dy, dyo, prec, G, bias_done=False, nxt=(None, 1.0)):
    x, n, pre, g, c, taps, bn, (p_res, sid_r) = saved
    d = x.shape[-1]
    dyb, dyo, fused_ok = _drop_grad(dy, dyo, p_res, sid_r, prec)
    ops.linear_wgrad(dyo, c, prec, _as2d(G.buf(conv.pointwise_conv2.weight)), side=True)
    if not (bias_done and fused_ok):
        ops.colsum_acc(dyb, G.buf(conv.pointwise_conv2.bias))
    dc = ops.linear(dyo, _wT(conv, "pw2", prec, (conv.pointwise_conv2.weight,)), None, prec=prec, out_dtype=act_dtype(prec))
    dtaps = torch.zeros_like(taps)
    if bn is not None:
        z, mean, var = bn
        if mean is None:            # GroupNorm (channel pairs, per frame)
            dz = ops.gn2_swish_bwd(z, dc, conv.norm.weight, conv.norm.bias, conv.norm.eps, G.buf(conv.norm.weight),
                                   G.buf(conv.norm.bias))
        else:
            dz, sums = ops.bn_swish_bwd(z, dc, mean, var, conv.norm.weight, conv.norm.bias, conv.norm.eps)
            G.buf(conv.norm.bias).add_(sums[0])
            G.buf(conv.norm.weight).add_(sums[1])
        dg = ops.dwconv_bwd(g, taps, dz, dtaps, G.buf(conv.depthwise_conv.bias), causal=conv.causal)
    else:
        dg = ops.conformer_conv_bwd(g, taps, conv.depthwise_conv.bias, conv.norm.weight, conv.norm.bias, conv.norm.eps, dc,
                                    dtaps, G.buf(conv.depthwise_conv.bias), G.buf(conv.norm.weight), G.buf(conv.norm.bias),
                                    causal=conv.causal)
    G.put(conv.depthwise_conv.weight, dtaps.t())
    if _fusable(pre) and pre.shape[-1] % 32 == 0:
        dpre = ops.act_bwd_bias(dg, pre, None, G.buf(conv.pointwise_conv1.bias), glu=True)
    else:
        dpre = ops.glu_bwd(dg, pre)
        ops.colsum_acc(dpre, G.buf(conv.pointwise_conv1.bias))
    ops.linear_wgrad(dpre, n, prec, _as2d(G.buf(conv.pointwise_conv1.weight)), side=True)
    dn = ops.linear(dpre, _wT(conv, "pw1", prec, (conv.pointwise_conv1.weight,)), None, prec=prec, out_dtype=torch.float32)
    return _ln_bwd(norm, dn, x, dy, G, prec, nxt[0], nxt[1])


# ------------------------------------------------------------------------------------------------
# encoder blocks
# ------------------------------------------------------------------------------------------------
def _param_list(block, extra):
    """Parameters of a block, with w_query / w_key / w_value weights moved next to each other (in that order) so that
    the fused QKV weight gradient is one contiguous region of the node's flat gradient buffer."""
    att = block.self_attn
    qkv = [att.w_query.weight, att.w_key.weight, att.w_value.weight]
    ids = {id(p) for p in qkv}
    rest = [p for p in block.parameters() if id(p) not in ids]
    return qkv + rest + [p for p in extra if p is not None]


class _BlockFn(torch.autograd.Function):
    """One encoder block (Conformer or Transformer) as a single autograd node."""

    @staticmethod
    def forward(ctx, xs, block, klens, pos, rel_bias, mask_kw, prec, in_scale, *params):
        u_bias, v_bias = rel_bias
        saved = {}
        x = xs
        if in_scale != 1.0:
            x = ops.scale_(xs.clone(), in_scale)
        pd = saved["p"] = float(block.dropout.p)          # dropout on every residual-branch output (block is in train mode)
        if getattr(block, "v2", False):           # conformer_block_v2.py: FFN -> conv -> plain MHA -> FFN -> LN
            x, saved["ffm"] = ffn_fwd(block.feed_forward_macaron, block.norm1, x, block.fc_factor, prec, pd)
            x, saved["conv"] = convmod_fwd(block.conv, block.norm2, x, prec, pd)
            x, saved["att"] = attn_fwd(block.self_attn, block.norm3, x, None, klens, None, None, mask_kw, prec, False, pd)
            x, saved["ff"] = ffn_fwd(block.feed_forward, block.norm4, x, block.fc_factor, prec, pd)
            saved["x5"] = x
            x = ops.layernorm(x, block.norm5.weight, block.norm5.bias, block.norm5.eps)
        elif hasattr(block, "feed_forward_macaron"):
            x, saved["ffm"] = ffn_fwd(block.feed_forward_macaron, block.norm1, x, block.fc_factor, prec, pd)
            x, saved["att"] = attn_fwd(block.self_attn, block.norm2, x, pos, klens, u_bias, v_bias, mask_kw, prec, True, pd)
            x, saved["conv"] = convmod_fwd(block.conv, block.norm3, x, prec, pd)
            x, saved["ff"] = ffn_fwd(block.feed_forward, block.norm4, x, block.fc_factor, prec, pd)
            saved["x5"] = x
            x = ops.layernorm(x, block.norm5.weight, block.norm5.bias, block.norm5.eps)
        else:
            x, saved["att"] = attn_fwd(block.self_attn, block.norm1, x, pos, klens, u_bias, v_bias, mask_kw, prec,
                                       block.rel_attn, pd)
            x, saved["ff"] = ffn_fwd(block.feed_forward, block.norm2, x, 1.0, prec, pd)
        ctx.block, ctx.saved, ctx.klens, ctx.pos, ctx.rel_bias = block, saved, klens, pos, rel_bias
        ctx.mask_kw, ctx.prec, ctx.in_scale, ctx.params = mask_kw, prec, in_scale, params
        return x

    @staticmethod
    def backward(ctx, dy):
        block, S, prec = ctx.block, ctx.saved, ctx.prec
        u_bias, v_bias = ctx.rel_bias
        G = _Grads(ctx.params)
        dy = dy.contiguous().float()
        # every LayerNorm backward also accumulates the bias gradient of the branch that consumes its dx -- valid only
        # without dropout on that branch's output (with dropout the branch sums its own masked gradient)
        nx = (lambda b, a: (None, 1.0)) if S["p"] > 0 else (lambda b, a: (b, a))
        if getattr(block, "v2", False):
            fc = block.fc_factor
            dx, dxo = _ln_bwd(block.norm5, dy, S["x5"], None, G, prec, *nx(block.feed_forward.out_bias, fc))
            dx, dxo = ffn_bwd(block.feed_forward, block.norm4, S["ff"], dx, dxo, fc, prec, G, bias_done=True,
                              nxt=nx(block.self_attn.w_out.bias, 1.0))
            dx, dxo = attn_bwd(block.self_attn, block.norm3, S["att"], dx, dxo, None, ctx.klens, None, None, ctx.mask_kw,
                               prec, False, G, (None, None), bias_done=True, nxt=nx(block.conv.pointwise_conv2.bias, 1.0))
            dx, dxo = convmod_bwd(block.conv, block.norm2, S["conv"], dx, dxo, prec, G, bias_done=True,
                                  nxt=nx(block.feed_forward_macaron.out_bias, fc))
            dx, dxo = ffn_bwd(block.feed_forward_macaron, block.norm1, S["ffm"], dx, dxo, fc, prec, G, bias_done=True)
        elif hasattr(block, "feed_forward_macaron"):
            fc = block.fc_factor
            dx, dxo = _ln_bwd(block.norm5, dy, S["x5"], None, G, prec, *nx(block.feed_forward.out_bias, fc))
            dx, dxo = ffn_bwd(block.feed_forward, block.norm4, S["ff"], dx, dxo, fc, prec, G, bias_done=True,
                              nxt=nx(block.conv.pointwise_conv2.bias, 1.0))
            dx, dxo = convmod_bwd(block.conv, block.norm3, S["conv"], dx, dxo, prec, G, bias_done=True,
                                  nxt=nx(block.self_attn.w_out.bias, 1.0))
            dx, dxo = attn_bwd(block.self_attn, block.norm2, S["att"], dx, dxo, ctx.pos, ctx.klens, u_bias, v_bias,
                               ctx.mask_kw, prec, True, G, ctx.rel_bias, bias_done=True,
                               nxt=nx(block.feed_forward_macaron.out_bias, fc))
            dx, dxo = ffn_bwd(block.feed_forward_macaron, block.norm1, S["ffm"], dx, dxo, fc, prec, G, bias_done=True)
        else:
            dx, dxo = ffn_bwd(block.feed_forward, block.norm2, S["ff"], dy, _gop(dy, prec), 1.0, prec, G,
                              nxt=nx(block.self_attn.w_out.bias, 1.0))
            dx, dxo = attn_bwd(block.self_attn, block.norm1, S["att"], dx, dxo, ctx.pos, ctx.klens, u_bias, v_bias,
                               ctx.mask_kw, prec, block.rel_attn, G, ctx.rel_bias, bias_done=True)
        if ctx.in_scale != 1.0:
            ops.scale_(dx, ctx.in_scale)
        ctx.saved = None
        G.done()
        grads = tuple(G.get(p) for p in ctx.params)
        return (dx, None, None, None, None, None, None, None) + grads


def block_forward(block, xs, klens, pos, rel_bias, mask_kw, prec, in_scale=1.0):
    """Training forward of one encoder block through its autograd node."""
    extra = rel_bias if (rel_bias[0] is not None and getattr(block.self_attn, "xl_like", False)) else ()
    params = _param_list(block, extra)
    rb = rel_bias if extra else (None, None)
    return _BlockFn.apply(xs, block, klens, pos, rb, mask_kw or {}, prec, float(in_scale), *params)


# ------------------------------------------------------------------------------------------------
# small nodes: LayerNorm, time max-pool, scaling
# ------------------------------------------------------------------------------------------------
class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        ctx.save_for_backward(x, weight)
        ctx.eps = eps
        return ops.layernorm(x, weight, bias, eps)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        flat = torch.zeros(2 * weight.numel(), dtype=torch.float32, device=weight.device)
        dg, db = flat[:weight.numel()], flat[weight.numel():]
        dx = ops.layernorm_bwd(dy.contiguous().float(), x, weight, ctx.eps, dres=None, dgamma=dg, dbeta=db)
        if _GRAD_SYNC is not None:
            _GRAD_SYNC(flat)
        return dx, dg, db, None


def layernorm(norm, x):
    return _LayerNormFn.apply(x, norm.weight, norm.bias, norm.eps)


class _MaxPoolTimeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, factor):
        ctx.save_for_backward(x)
        ctx.factor = factor
        return ops.pool_time(x, factor, "max")

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.maxpool_time_bwd(x, dy, ctx.factor), None


def maxpool_time(x, factor):
    return _MaxPoolTimeFn.apply(x, factor)


class _PoolTimeFn(torch.autograd.Function):
    """mean / drop / add time pooling (MeanPool / Drop / Add subsamplers) with the CUDA backward."""

    @staticmethod
    def forward(ctx, x, factor, mode):
        ctx.T, ctx.factor, ctx.mode = x.shape[1], factor, mode
        return ops.pool_time(x, factor, mode)

    @staticmethod
    def backward(ctx, dy):
        return ops.pool_time_bwd(dy, ctx.T, ctx.factor, ctx.mode), None, None


def subsample_train(sub, xs, xlens):
    """Training forward of one intermediate subsampler (encoders/subsampling.py) -> (xs, xlens): the pooling variants are
    one autograd node each; concat / conv1d are GEMMs with a fused ReLU over re-laid-out frames (data movement only)."""
    from .encoders.subsampling import ConcatSubsampler, Conv1dSubsampler
    from .modules._prep import get_precision
    f = sub.factor
    if f == 1:
        return xs, xlens
    if isinstance(sub, ConcatSubsampler):
        B, T, D = xs.shape
        x = xs[:, :(T // f) * f].reshape(B, T // f, f * D)
        return linear_relu(sub, 'proj', sub.proj.weight, sub.proj.bias, x, get_precision(sub)), sub._lens(xlens)
    if isinstance(sub, Conv1dSubsampler):
        B, T, D = xs.shape
        k, pad = sub.kernel_size, (sub.kernel_size - 1) // 2
        To = (T + 2 * pad - (k - 1) - 1) // f + 1
        xp = torch.nn.functional.pad(xs, (0, 0, pad, pad))
        cols = xp.unfold(1, k, f)[:, :To].reshape(B, To, D * k)             # column c * k + j = nn.Conv1d weight order
        return (linear_relu(sub, 'conv1d_ck', sub.conv1d.weight, sub.conv1d.bias, cols, get_precision(sub)),
                sub._lens(xlens))
    if sub.mode == "max":
        return maxpool_time(xs, f), sub._lens(xlens)
    return _PoolTimeFn.apply(xs, f, sub.mode), sub._lens(xlens)


class _ScaleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, a):
        ctx.a = a
        return ops.scale_(x.contiguous().clone(), a)

    @staticmethod
    def backward(ctx, dy):
        return ops.scale_(dy.contiguous().clone(), ctx.a), None


def scale(x, a):
    return x if a == 1.0 else _ScaleFn.apply(x, float(a))


class _DropoutFn(torch.autograd.Function):
    """nn.Dropout with the library's kernel: the backward regenerates the mask from the call's stream id."""

    @staticmethod
    def forward(ctx, x, p):
        ctx.p, ctx.sid = p, nrandom.next_stream()
        return ops.dropout(x.detach(), p, ctx.sid)

    @staticmethod
    def backward(ctx, dy):
        return ops.dropout(dy.contiguous(), ctx.p, ctx.sid), None


def dropout(x, p):
    """Training-mode dropout of an fp32 / bf16 CUDA tensor (identity for p == 0)."""
    return x if p <= 0 else _DropoutFn.apply(x, float(p))


class _AddPosEncFn(torch.autograd.Function):
    """y = x * a + pe[offset:offset+T]  (PositionalEncoding 'add' / 'none'); dx = a * dy."""

    @staticmethod
    def forward(ctx, x, pos_enc, offset):
        ctx.a = pos_enc.scale
        return pos_enc(x.detach().clone(), scale=True, offset=offset)

    @staticmethod
    def backward(ctx, dy):
        return ops.scale_(dy.contiguous().float().clone(), ctx.a), None, None


def add_pos_enc(pos_enc, x, offset=0):
    return _AddPosEncFn.apply(x, pos_enc, offset)


def linear(module, name, lin, x, prec):
    """nn.Linear through the tcgen05 GEMM with autograd (dgrad + wgrad on the same kernels)."""
    from .modules.linear import _LinearFn
    return _LinearFn.apply(x, lin.weight, lin.bias, lin, prec)


class _LinearReluFn(torch.autograd.Function):
    """relu(x W^T + b) with the ReLU fused into the GEMM epilogue; backward masks dy with the saved output."""

    @staticmethod
    def forward(ctx, x, weight, bias, owner, name, prec):
        y = ops.linear(x, prepared(owner, name, prec, (weight,), build=_as2d), bias, prec=prec, act="relu", out_dtype=torch.float32)
        ctx.save_for_backward(x, weight, y)
        ctx.owner, ctx.name, ctx.prec = owner, name, prec
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        prec = ctx.prec
        dz = ops.relu_mask(dy.contiguous().float(), y).reshape(-1, y.shape[-1])
        dzo = _gop(dz, prec)
        N = weight.shape[0]
        flat = torch.zeros(weight.numel() + N, dtype=torch.float32, device=weight.device)
        gw, gb = flat[:weight.numel()].view(N, -1), flat[weight.numel():]
        dx = ops.linear(dzo, _wT(ctx.owner, ctx.name, prec, (weight,)), None, prec=prec, out_dtype=torch.float32).reshape(x.shape)
        ops.linear_wgrad(dzo, x.reshape(-1, x.shape[-1]).float(), prec, gw)
        ops.colsum_acc(dz, gb)
        if _GRAD_SYNC is not None:
            _GRAD_SYNC(flat)
        return dx, gw.view(weight.shape), gb, None, None, None


def linear_relu(owner, name, weight, bias, x, prec):
    """ReLU(Linear) with autograd; `weight` may be a Conv1d weight [N, C, k] whose GEMM view the caller has laid out in x."""
    return _LinearReluFn.apply(x, weight, bias, owner, name, prec)


# ------------------------------------------------------------------------------------------------
# one (bi)directional LSTM layer as one node (reference: autograd through Padding / nn.LSTM, encoders/rnn.py:534-546)
# ------------------------------------------------------------------------------------------------
class _LstmLayerFn(torch.autograd.Function):
    """params = (w_ih, w_hh, b_ih, b_hh) of the forward direction [+ the same four of the reverse direction]."""

    @staticmethod
    def forward(ctx, xs, enc, lth, lens_dev, prec, *params):
        nd = len(params) // 4
        w_ih, w_hh = params[0::4], params[1::4]
        b_ih, b_hh = params[2::4], params[3::4]
        w_ihp = prepared(enc, 'w_ih%d' % lth, prec, tuple(w_ih), build=lambda *ws: torch.cat(ws, dim=0))
        bias = cached(enc, 'b%d' % lth, tuple(b_ih) + tuple(b_hh),
                      lambda *bs: (torch.cat(bs[:nd]) + torch.cat(bs[nd:])).float().contiguous())
        whh = cached(enc, 'w_hh%d' % lth, tuple(w_hh), lambda *ws: torch.stack(ws, dim=0).float().contiguous())
        xs = xs.contiguous().float()
        gates_x = ops.linear(xs, w_ihp, bias, prec=prec, out_dtype=torch.float32)
        ys, acts, cprev, hprev = ops.lstm_seq(gates_x, whh, lens_dev, nd, save=True, prec=prec)
        ctx.save_for_backward(xs, acts, cprev, hprev, whh, lens_dev)
        ctx.enc, ctx.lth, ctx.prec, ctx.params, ctx.nd = enc, lth, prec, params, nd
        return ys

    @staticmethod
    def backward(ctx, dy):
        xs, acts, cprev, hprev, whh, lens_dev = ctx.saved_tensors
        prec, nd, params = ctx.prec, ctx.nd, ctx.params
        w_ih, w_hh = params[0::4], params[1::4]
        b_ih, b_hh = params[2::4], params[3::4]
        H4 = acts.shape[-1]
        # bucket order: [w_ih of all directions | w_hh | b_ih | b_hh]: the direction-concatenated gradients are single views
        G = _Grads(tuple(w_ih) + tuple(w_hh) + tuple(b_ih) + tuple(b_hh))
        dg = ops.lstm_seq_bwd(dy, acts, cprev, whh, lens_dev, prec=prec)                     # [B, T, nd*4H] fp32
        dgo = _gop(dg, prec)
        dx = None
        if ctx.needs_input_grad[0]:
            w_ihT = _wT(ctx.enc, 'w_ih%d' % ctx.lth, prec, tuple(w_ih), build=lambda *ws: torch.cat(ws, dim=0))
            dx = ops.linear(dgo, w_ihT, None, prec=prec, out_dtype=torch.float32)
        ops.linear_wgrad(dgo, xs, prec, G.fused(w_ih))                            # all directions at once
        for d in range(nd):
            ops.linear_wgrad(dgo[:, :, d * H4:(d + 1) * H4].contiguous(), hprev[:, :, d].contiguous(), prec, G.get(w_hh[d]))
        gb = G.fused(b_ih)
        ops.colsum_acc(dg, gb.view(-1))
        G.fused(b_hh).copy_(gb)                                                    # b_ih and b_hh enter the gates as a sum
        G.done()
        return (dx, None, None, None, None) + tuple(G.get(p) for p in params)


def lstm_layer(enc, lth, xs, lens_dev, prec):
    rnn = enc.rnn[lth]
    params = []
    for sfx in ['_l0'] + (['_l0_reverse'] if enc.bidirectional else []):
        params += [getattr(rnn, 'weight_ih' + sfx), getattr(rnn, 'weight_hh' + sfx),
                   getattr(rnn, 'bias_ih' + sfx), getattr(rnn, 'bias_hh' + sfx)]
    return _LstmLayerFn.apply(xs, enc, lth, lens_dev, prec, *params)


class _LstmChunkFn(torch.autograd.Function):
    """One UNIDIRECTIONAL LSTM layer over a chunk with a carried state (latency-controlled BLSTM training, rnn.py:454-498):
    (xs, h0, c0) -> (ys, hN, cN); h0 / c0 may be None (zero state).  The backward receives the gradient w.r.t. the final state
    from the next chunk's node and returns the gradient w.r.t. the initial one (persistent BPTT kernel, state arguments)."""

    @staticmethod
    def forward(ctx, xs, h0, c0, owner, tag, lens_dev, prec, w_ih, w_hh, b_ih, b_hh):
        w_ihp = prepared(owner, 'w_ih' + tag, prec, (w_ih,))
        bias = cached(owner, 'b' + tag, (b_ih, b_hh), lambda a, b: (a + b).float().contiguous())
        whh = cached(owner, 'w_hh' + tag, (w_hh,), lambda w: w.unsqueeze(0).float().contiguous())
        xs = xs.contiguous().float()
        gates_x = ops.linear(xs, w_ihp, bias, prec=prec, out_dtype=torch.float32)
        state = (h0.detach(), c0.detach()) if h0 is not None else None
        ys, acts, cprev, hprev, (hN, cN) = ops.lstm_seq(gates_x, whh, lens_dev, 1, save=True, state=state, want_state=True, prec=prec)
        ctx.save_for_backward(xs, acts, cprev, hprev, whh, lens_dev)
        ctx.owner, ctx.tag, ctx.prec, ctx.params, ctx.has_state = owner, tag, prec, (w_ih, w_hh, b_ih, b_hh), h0 is not None
        return ys, hN, cN

    @staticmethod
    def backward(ctx, dy, dhN, dcN):
        xs, acts, cprev, hprev, whh, lens_dev = ctx.saved_tensors
        prec = ctx.prec
        w_ih, w_hh, b_ih, b_hh = ctx.params
        G = _Grads((w_ih, w_hh, b_ih, b_hh))
        dstate = None
        if dhN is not None or dcN is not None:
            z = torch.zeros_like(hprev[:, 0].transpose(0, 1))            # [1, B, H]
            dstate = (dhN if dhN is not None else z, dcN if dcN is not None else z)
        if dy is None:
            dy = torch.zeros(xs.shape[0], xs.shape[1], whh.shape[-1], dtype=torch.float32, device=xs.device)
        out = ops.lstm_seq_bwd(dy, acts, cprev, whh, lens_dev, dstate=dstate, want_dstate=ctx.has_state, prec=prec)
        dg, d0 = out if ctx.has_state else (out, (None, None))
        dgo = _gop(dg, prec)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.linear(dgo, _wT(ctx.owner, 'w_ih' + ctx.tag, prec, (w_ih,)), None, prec=prec, out_dtype=torch.float32)
        ops.linear_wgrad(dgo, xs, prec, G.buf(w_ih))
        ops.linear_wgrad(dgo, hprev[:, :, 0].contiguous(), prec, G.buf(w_hh))
        ops.colsum_acc(dg, G.buf(b_ih))
        G.buf(b_hh).copy_(G.buf(b_ih))
        G.done()
        return dx, d0[0], d0[1], None, None, None, None, G.get(w_ih), G.get(w_hh), G.get(b_ih), G.get(b_hh)


def lstm_chunk(owner, tag, rnn, xs, lens_dev, state, prec):
    """-> (ys `[B,T,H]`, (hN, cN) `[1,B,H]` each, attached to the graph)."""
    h0, c0 = state if state is not None else (None, None)
    ys, hN, cN = _LstmChunkFn.apply(xs, h0, c0, owner, tag, lens_dev, prec, rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0,
                                    rnn.bias_hh_l0)
    return ys, (hN, cN)


# ------------------------------------------------------------------------------------------------
# RNN-T joint network + loss as one node
# ------------------------------------------------------------------------------------------------
class _RnntJointLossFn(torch.autograd.Function):
    """(w_enc(e) `[B,T,J]`, w_dec(d) `[B,U+1,J]`) -> mean RNN-T loss (reference decoders/rnn_transducer.py:240-252, :262-276):
    tanh of the broadcast sum, output layer, log-softmax and the lattice in ONE autograd node.  Forward keeps the joint
    activation h, the log-probabilities and the small lattice workspace (alpha, beta, gathered emissions); backward = one pass
    that turns log-probs into d loss / d logits (log-softmax backward folded into the lattice gradient, scaled by the upstream
    gradient on the device, written directly in the GEMM operand dtype -- d loss / d log_probs is never materialised), the
    output layer's dgrad / wgrad GEMMs and the two broadcast reductions of the tanh backward."""

    @staticmethod
    def forward(ctx, e, d, weight, bias, owner, labels, flens, ylens, blank, prec):
        B, T, J = e.shape
        U1, V = d.shape[1], weight.shape[0]
        h = ops.rnnt_joint_tanh(e.detach(), d.detach(), out_dtype=act_dtype(prec))
        logits = ops.linear(h.view(B * T * U1, J), prepared(owner, "output", prec, (weight,)), bias, prec=prec,
                            out_dtype=torch.float32)
        lp = ops.softmax_rows(logits.view(B, T, U1, V), log=True, inplace=True)
        loss, nll, _, ws = ops.rnnt_loss_fwd_bwd(lp, labels, flens, ylens, blank, need_grad=False, return_ws=True)
        ctx.keep = (h, lp, ws, nll, labels, flens, ylens, blank, weight, bias)
        ctx.owner, ctx.prec = owner, prec
        ctx.mark_non_differentiable(nll)
        return loss, nll

    @staticmethod
    def backward(ctx, g_loss, g_nll):
        h, lp, ws, nll, labels, flens, ylens, blank, weight, bias = ctx.keep
        ctx.keep = None
        prec = ctx.prec
        B, T, U1, J = h.shape
        V = weight.shape[0]
        bf16 = prec == "bf16"
        dz = ops.rnnt_grad_logits(lp, ws, nll, labels, flens, ylens, blank, g_loss.detach().float(),
                                  out_dtype=torch.bfloat16 if bf16 else torch.float32, inplace=not bf16).view(-1, V)
        dzo = dz
        nb = V if bias is not None else 0
        flat = torch.zeros(weight.numel() + nb, dtype=torch.float32, device=weight.device)
        gw = flat[:weight.numel()].view(weight.shape)
        gb = flat[weight.numel():] if bias is not None else None
        ops.linear_wgrad(dzo, h.view(-1, J), prec, gw)
        if gb is not None:
            ops.colsum_acc(dz, gb)
        dh = ops.linear(dzo, _wT(ctx.owner, "output", prec, (weight,)), None, prec=prec, out_dtype=act_dtype(prec))
        de, dd = ops.rnnt_joint_tanh_bwd(h, dh.view(B, T, U1, J))
        if _GRAD_SYNC is not None:
            _GRAD_SYNC(flat)
        return de, dd, gw, gb, None, None, None, None, None, None


def rnnt_joint_loss(owner, e, d, labels, flens, ylens, blank, prec):
    """-> (loss 0-dim, nll [B]); `owner.output` is the vocabulary layer."""
    return _RnntJointLossFn.apply(e, d, owner.output.weight, owner.output.bias, owner, labels, flens, ylens, blank, prec)


# ------------------------------------------------------------------------------------------------
# CNN front-end (Conv2dBlock stack + bridge) as one node
# ------------------------------------------------------------------------------------------------
def _conv_any(block, name, layer, x, B, T, F, first, prec, weight=None, relu=True):
    """3x3 conv (+bias, +ReLU): tcgen05 implicit GEMM for 32->32 bf16, CUDA-core kernel otherwise."""
    ci, co = layer.in_channels, layer.out_channels
    adt = act_dtype(prec)
    if prec == "bf16" and ci == 32 and co == 32 and x.dtype == torch.bfloat16:
        if weight is None:
            wt = prepared(block, name + ".taps", "bf16", (layer.weight,),
                          build=lambda w: w.permute(0, 2, 3, 1).reshape(32, 288))[0][:, :288].contiguous()
            bias = layer.bias
        else:
            wt = prepared(block, name + ".dtaps", "bf16", (layer.weight,),
                          build=lambda w: ops.conv3x3_dgrad_weight(w).permute(0, 2, 3, 1).reshape(32, 288))[0][:, :288].contiguous()
            bias = cached(block, name + ".zero_bias", (layer.bias,), lambda b: torch.zeros_like(b, dtype=torch.float32))
        return ops.conv3x3_c32_tc(x.view(B, T, F, 32), wt, bias, relu=relu, pool2x2=False)
    if weight is None:
        return ops.conv3x3_relu(x, layer.weight, layer.bias, B, T, F, in_chmajor=first, relu=relu, out_dtype=adt)
    wd = cached(block, name + ".dgrad_w", (layer.weight,), lambda w: ops.conv3x3_dgrad_weight(w).float())
    zb = cached(block, name + ".zero_bias_in", (layer.weight,), lambda w: torch.zeros(w.shape[1], dtype=torch.float32, device=w.device))
    return ops.conv3x3_relu(x, wd, zb, B, T, F, in_chmajor=False, relu=False, out_dtype=adt)


class _FrontendFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xs, enc, out_scale, prec, *params):
        B, T, Fdim = xs.shape
        F = Fdim // enc.in_channel
        x = xs.contiguous().float()
        tape = []
        n = len(enc.layers)
        for i, blk in enumerate(enc.layers):
            last_chmajor = (i == n - 1) and enc.bridge is None
            pt, pf = blk.pooling if blk.pool is not None else (1, 1)
            st, sf = getattr(blk, "stride", (1, 1))
            general = blk.norm1 is not None or getattr(blk, "residual_active", False)
            if general:
                # normalised and / or residual block (conv.py:360-394), fp32 channels-last:
                #   conv1 -> [norm1] -> ReLU -> conv2(stride) -> [norm2] -> [+ block input] -> ReLU -> [pool]
                xin = x if i == 0 else x.float()
                res = xin.view(B, T, F, -1) if getattr(blk, "residual_active", False) else None
                rec = dict(x=xin, T=T, F=F, pt=pt, pf=pf, chmajor=last_chmajor, first=(i == 0), general=True, res=res is not None)
                a1, rec["z1"], rec["st1"] = _norm_stage(blk, "conv1", blk.conv1, blk.norm1, xin, B, T, F, i == 0, (1, 1), None)
                a2s, rec["z2"], rec["st2"] = _norm_stage(blk, "conv2", blk.conv2, blk.norm2, a1, B, T, F, False, (st, sf), res)
                rec.update(a1=a1, a2=a2s, a2s=a2s)
                if (st, sf) != (1, 1):
                    rec.update(stride=(st, sf), Ts=a2s.size(1), Fs=a2s.size(2))
                    T, F = a2s.size(1), a2s.size(2)
            else:
                a1 = _conv_any(blk, "conv1", blk.conv1, x, B, T, F, i == 0, prec)
                a2 = _conv_any(blk, "conv2", blk.conv2, a1, B, T, F, False, prec)
                rec = dict(x=x, a1=a1, a2=a2, T=T, F=F, pt=pt, pf=pf, chmajor=last_chmajor, first=(i == 0))
                a2s = a2
                if (st, sf) != (1, 1):      # strided 'same' conv = the stride-1 conv sampled at 0, s, 2s, ... (ReLU commutes)
                    a2s = a2.view(B, T, F, -1)[:, ::st, ::sf].contiguous()
                    rec.update(stride=(st, sf), a2s=a2s, Ts=a2s.size(1), Fs=a2s.size(2))
                    T, F = a2s.size(1), a2s.size(2)
            if blk.pool is not None or last_chmajor:
                x = ops.maxpool2d(a2s.view(B, T, F, -1), pt, pf, out_chmajor=last_chmajor)
                rec["pooled"] = True
                T, F = -(-T // pt), -(-F // pf)
            else:
                x = a2s
                rec["pooled"] = False
            tape.append(rec)
        C, Fo = enc._c_last, F
        if enc.bridge is not None:
            wb = prepared(enc, "bridge", prec, (enc.bridge.weight,),
                          build=lambda w: w.view(w.size(0), C, Fo).transpose(1, 2).reshape(w.size(0), Fo * C))
            feat = x.reshape(B, T, Fo * C)
            y = ops.linear(feat, wb, enc.bridge.bias, prec=prec, alpha=out_scale, out_dtype=torch.float32)
            ctx.feat = feat
        else:
            y = x.float() if x.dtype != torch.float32 else x
            y = y.reshape(B, T, C * Fo)
            if out_scale != 1.0:
                y = ops.scale_(y.contiguous().clone(), out_scale)
        ctx.enc, ctx.tape, ctx.prec, ctx.out_scale, ctx.params = enc, tape, prec, out_scale, params
        ctx.dims = (B, T, Fo, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        enc, tape, prec = ctx.enc, ctx.tape, ctx.prec
        B, Tl, Fo, C = ctx.dims
        G = _Grads(ctx.params)
        adt = act_dtype(prec)
        dy = dy.contiguous().float()
        if enc.bridge is not None:
            dyo = _gop(dy, prec)
            gperm = torch.zeros(enc.bridge.weight.shape[0], Fo * C, dtype=torch.float32, device=dy.device)
            ops.linear_wgrad(dyo, ctx.feat, prec, gperm, alpha=ctx.out_scale)
            # undo the (f, c) -> (c, f) column permutation of the cached operand: pure view + copy by autograd's accumulate
            G.put(enc.bridge.weight, gperm.view(-1, Fo, C).transpose(1, 2))
            ops.colsum_acc(dy, G.buf(enc.bridge.bias), alpha=ctx.out_scale)
            wbT = prepared(enc, "bridge.T", prec, (enc.bridge.weight,),
                           build=lambda w: w.view(w.size(0), C, Fo).transpose(1, 2).reshape(w.size(0), Fo * C).t().contiguous())
            d = ops.linear(dyo, wbT, None, prec=prec, alpha=ctx.out_scale, out_dtype=adt).view(B, Tl, Fo, C)
        else:
            d = dy if ctx.out_scale == 1.0 else ops.scale_(dy.clone(), ctx.out_scale)
            if adt == torch.bfloat16:
                d = ops.to_bf16(d)
        for blk, rec in zip(reversed(list(enc.layers)), reversed(tape)):
            T, F = rec["T"], rec["F"]
            a1, a2, x = rec["a1"], rec["a2"], rec["x"]
            if rec.get("general"):
                d = _general_block_bwd(blk, rec, d, G, B)
                continue
            if "stride" in rec:             # gradient of the sampled positions, scattered back onto the stride-1 grid
                a2s, Ts, Fs = rec["a2s"], rec["Ts"], rec["Fs"]
                if rec["pooled"]:
                    dzs = ops.maxpool2d_relu_bwd(a2s.view(B, Ts, Fs, -1), d, rec["pt"], rec["pf"], in_chmajor=rec["chmajor"])
                else:
                    dzs = ops.relu_mask(d.reshape(a2s.shape).to(a2s.dtype), a2s)
                dz2 = torch.zeros(B, T, F, a2s.shape[-1], dtype=dzs.dtype, device=dzs.device)
                dz2[:, ::rec["stride"][0], ::rec["stride"][1]] = dzs.view(B, Ts, Fs, -1)
            elif rec["pooled"]:
                dz2 = ops.maxpool2d_relu_bwd(a2.view(B, T, F, -1), d, rec["pt"], rec["pf"], in_chmajor=rec["chmajor"])
            else:
                dz2 = ops.relu_mask(d.reshape(a2.shape).to(a2.dtype), a2)
            ops.conv3x3_wgrad(a1, dz2, G.buf(blk.conv2.weight), G.buf(blk.conv2.bias), B, T, F)
            da1 = _conv_any(blk, "conv2", blk.conv2, dz2, B, T, F, False, prec, weight="dgrad", relu=False)
            dz1 = ops.relu_mask(da1.reshape(a1.shape), a1)
            ops.conv3x3_wgrad(x, dz1, G.buf(blk.conv1.weight), G.buf(blk.conv1.bias), B, T, F, in_chmajor=rec["first"])
            if not rec["first"]:
                d = _conv_any(blk, "conv1", blk.conv1, dz1, B, T, F, False, prec, weight="dgrad", relu=False)
        ctx.tape = None
        G.done()
        return (None, None, None, None) + tuple(G.get(p) for p in ctx.params)


def _ln2d_affine(blk, name, norm):
    """LayerNorm2D's [C, F] affine parameters in the (f, c) order of a channels-last frame."""
    gam = cached(blk, name + ".ln_w", (norm.norm.weight,), lambda t: t.t().contiguous().reshape(-1).float())
    bet = cached(blk, name + ".ln_b", (norm.norm.bias,), lambda t: t.t().contiguous().reshape(-1).float())
    return gam, bet


def _norm_stage(blk, name, conv, norm, x, B, T, F, first, stride, res):
    """One conv -> [norm] -> [+ res] -> ReLU stage of a general CNN block in training, fp32 channels-last.
    -> (a, z, stat): a = the rectified output, z = what the norm's backward needs (the conv output), stat = BatchNorm's batch
    (mean, var).

    LayerNorm2D: the library's LayerNorm over a frame's F*C values with the affine parameters in (f, c) order.
    BatchNorm2d (train() mode): per-channel (sum, sum of squares) over all B*T*F positions from the k = 1 case of the statistics
    kernel; running statistics updated as nn.BatchNorm2d does; the normalised output is the SAME conv kernel run with the batch
    statistics folded into its weights (with its fused ReLU when no residual is added in between)."""
    C = conv.out_channels

    def sample(t):
        t = t.view(B, T, F, C)
        return t if stride == (1, 1) else t[:, ::stride[0], ::stride[1]].contiguous()
    stat = None
    if norm is None:                                        # residual block without normalisation
        z = None
        n = sample(ops.conv3x3_relu(x, conv.weight, conv.bias, B, T, F, in_chmajor=first, relu=res is None, out_dtype=torch.float32))
        if res is None:
            return n, None, None
    elif isinstance(norm, torch.nn.BatchNorm2d):
        z = sample(ops.conv3x3_relu(x, conv.weight, conv.bias, B, T, F, in_chmajor=first, relu=False, out_dtype=torch.float32))
        M = z.numel() // C
        ones = cached(blk, name + ".bn_ones", (norm.weight,), lambda w: torch.ones(1, w.numel(), dtype=torch.float32, device=w.device))
        zero = cached(blk, name + ".bn_zero", (norm.weight,), lambda w: torch.zeros(w.numel(), dtype=torch.float32, device=w.device))
        _, stats = ops.dwconv_stats(z.view(1, M, C), ones, zero)
        stat = _bn_batch_stats(norm, stats, M)
        sc = norm.weight.detach().float() / torch.sqrt(stat[1] + norm.eps)           # C-length vectors
        wf = (conv.weight.detach().float() * sc.view(-1, 1, 1, 1)).contiguous()
        bf = ((conv.bias.detach().float() - stat[0]) * sc + norm.bias.detach().float()).contiguous()
        n = sample(ops.conv3x3_relu(x, wf, bf, B, T, F, in_chmajor=first, relu=res is None, out_dtype=torch.float32))
        if res is None:
            return n, z, stat
    else:
        z = sample(ops.conv3x3_relu(x, conv.weight, conv.bias, B, T, F, in_chmajor=first, relu=False, out_dtype=torch.float32))
        gam, bet = _ln2d_affine(blk, name, norm)
        n = ops.layernorm(z.reshape(z.shape[0] * z.shape[1], -1), gam, bet, norm.norm.eps).view(z.shape)
    if res is not None and res.shape == n.shape:
        n = ops.dropout_add(n.contiguous(), res.contiguous(), 0.0, 1.0, 0)           # p = 0: plain out = res + n
    return ops.relu_mask(n, n).view(n.shape), z, stat


def _general_block_bwd(blk, rec, d, G, B):
    """Backward of one normalised / residual block of the CNN front-end (fp32): pool / ReLU mask -> [residual branch] ->
    LayerNorm2D or BatchNorm2d backward -> conv gradients, twice."""
    T, F = rec["T"], rec["F"]
    a1, a2, x, z1, z2 = rec["a1"], rec["a2"], rec["x"], rec["z1"], rec["z2"]
    Ts, Fs = rec.get("Ts", T), rec.get("Fs", F)
    C = a2.shape[-1]

    def norm_bwd(dn, z, norm, stat, t, f):
        if norm is None:
            return dn.view(B, t, f, C)
        if isinstance(norm, torch.nn.BatchNorm2d):
            dz, sums = ops.bn_bwd(z.reshape(-1, C), dn.reshape(-1, C).float(), stat[0], stat[1],
                                  norm.weight.detach().float().contiguous(), norm.eps)
            G.put(norm.bias, sums[0])
            G.put(norm.weight, sums[1])
            return dz.view(B, t, f, C)
        gam, _ = _ln2d_affine(blk, "conv1" if norm is blk.norm1 else "conv2", norm)
        dgam = torch.zeros(f * C, dtype=torch.float32, device=z.device)
        dbet = torch.zeros(f * C, dtype=torch.float32, device=z.device)
        dz = ops.layernorm_bwd(dn.reshape(B * t, f * C).float(), z.reshape(B * t, f * C), gam, norm.norm.eps, dgamma=dgam, dbeta=dbet)
        G.put(norm.norm.weight, dgam.view(f, C).t())            # back to the parameter's [C, F] layout (tiny)
        G.put(norm.norm.bias, dbet.view(f, C).t())
        return dz.view(B, t, f, C)

    if rec["pooled"]:
        dn2 = ops.maxpool2d_relu_bwd(a2.view(B, Ts, Fs, C), d.float() if d.dtype != torch.float32 else d, rec["pt"], rec["pf"],
                                     in_chmajor=rec["chmajor"])
    else:
        dn2 = ops.relu_mask(d.reshape(a2.shape).float(), a2)
    dn2 = dn2.view(B, Ts, Fs, C)
    d_res = dn2 if (rec["res"] and tuple(x.reshape(B, T, F, -1).shape) == tuple(dn2.shape)) else None
    dz2s = norm_bwd(dn2, z2, blk.norm2, rec["st2"], Ts, Fs)
    if "stride" in rec:
        dz2 = torch.zeros(B, T, F, C, dtype=torch.float32, device=dz2s.device)
        dz2[:, ::rec["stride"][0], ::rec["stride"][1]] = dz2s
    else:
        dz2 = dz2s.contiguous()
    ops.conv3x3_wgrad(a1, dz2, G.buf(blk.conv2.weight), G.buf(blk.conv2.bias), B, T, F)
    wd2 = cached(blk, "conv2.dgrad_w", (blk.conv2.weight,), lambda w: ops.conv3x3_dgrad_weight(w).float())
    zb2 = cached(blk, "conv2.zero_bias_in", (blk.conv2.weight,), lambda w: torch.zeros(w.shape[1], dtype=torch.float32, device=w.device))
    da1 = ops.conv3x3_relu(dz2, wd2, zb2, B, T, F, in_chmajor=False, relu=False, out_dtype=torch.float32)
    dn1 = ops.relu_mask(da1.reshape(a1.shape), a1)
    dz1 = norm_bwd(dn1, z1, blk.norm1, rec["st1"], T, F).contiguous()
    ops.conv3x3_wgrad(x, dz1, G.buf(blk.conv1.weight), G.buf(blk.conv1.bias), B, T, F, in_chmajor=rec["first"])
    if rec["first"]:
        return None
    wd1 = cached(blk, "conv1.dgrad_w", (blk.conv1.weight,), lambda w: ops.conv3x3_dgrad_weight(w).float())
    zb1 = cached(blk, "conv1.zero_bias_in", (blk.conv1.weight,), lambda w: torch.zeros(w.shape[1], dtype=torch.float32, device=w.device))
    dx = ops.conv3x3_relu(dz1, wd1, zb1, B, T, F, in_chmajor=False, relu=False, out_dtype=torch.float32)
    if d_res is not None:                                       # the skip connection's share of the block-input gradient
        dx = ops.dropout_add(dx.contiguous(), d_res.contiguous().view(dx.shape), 0.0, 1.0, 0)
    return dx


def frontend_check(enc):
    """Raise NotImplementedError for CNN front-end variants without a training path (inference handles them)."""
    if getattr(enc, "is_1dconv", False):
        raise NotImplementedError("1-D CNN front-end: inference only on the B200 path")
    for blk in enc.layers:
        if blk.training and blk.dropout.p > 0:
            raise NotImplementedError("dropout > 0 in the CNN front-end is not on the B200 path (build_encoder passes 0)")
        if not blk.trainable or (getattr(blk, "residual_active", False) and blk is enc.layers[0]):
            raise NotImplementedError("CNN front-end: residual connection around the first block (raw feature layout)")


def frontend_forward(enc, xs, out_scale, prec):
    frontend_check(enc)
    params = [p for p in enc.parameters()]
    return _FrontendFn.apply(xs, enc, float(out_scale), prec, *params)
