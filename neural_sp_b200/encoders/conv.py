"""VGG-style CNN front-end (reference encoders/conv.py:18-195, 289-396, 451-505), B200-native.

Same constructor, ``forward(xs, xlens, lookback, lookahead) -> (xs, xlens)`` and state_dict keys
(``layers.N.conv1/conv2.{weight,bias}``, ``bridge.{weight,bias}``).  Activations are channels-last
``[B, T, F, C]`` on the device; the reference's final flatten order (``c * F' + f``, conv.py:189) is
produced by the last pooling kernel (or folded into the bridge weight's column order)."""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..modules._prep import prepared, cached, get_precision, act_dtype
from ..modules.initialization import init_with_lecun_normal
from .encoder_base import EncoderBase


def parse_cnn_config(channels, kernel_sizes, strides, poolings):
    """'32_32', '(3,3)_(3,3)', ... -> lists (reference conv.py:480-505)."""
    def pairs(s):
        return [[int(v) for v in tok.strip('()').split(',')] for tok in s.split('_')] if s else []
    is_1dconv = '(' not in kernel_sizes
    if is_1dconv:
        to_i = lambda s: [int(c) for c in s.split('_')] if s else []   # noqa: E731
        return (to_i(channels), to_i(kernel_sizes), to_i(strides), to_i(poolings)), True
    chans = [int(c) for c in channels.split('_')] if channels else []
    return (chans, pairs(kernel_sizes), pairs(strides), pairs(poolings)), False


def _conv_len(n, stride):       # k=3, p=1: floor((n + 2 - 2 - 1)/s) + 1   (conv.py:476-477)
    return (n - 1) // stride + 1


def _pool_len(n, k):            # ceil-mode pool, kernel = stride = k      (conv.py:472-474)
    return (n + 1 - k) // k + 1


class LayerNorm2D(nn.Module):
    """Layer normalisation over (channel, frequency) of every frame (reference conv.py:399-421); parameter ``norm.weight
    [C, F]`` as in the reference.  On channels-last activations it is a LayerNorm over the F*C contiguous values of a frame with
    the affine parameters permuted to (f, c) order."""

    def __init__(self, channel, idim, eps=1e-12):
        super().__init__()
        self.norm = nn.LayerNorm([channel, idim], eps=eps)


class Conv2dBlock(EncoderBase):
    """conv3x3 -> [norm] -> ReLU -> conv3x3(stride) -> [norm] -> [+ residual] -> ReLU -> max-pool (reference conv.py:289-396).

    The recipes' shape (stride 1, no normalisation, no residual) runs on the fused kernels (tcgen05 implicit GEMM with ReLU /
    2x2 pooling in the epilogue in bf16 mode) and has a training path.  The general shape is inference-only and composed from
    the same conv kernel: a strided 'same' 3x3 conv is the stride-1 conv sampled every s-th position; BatchNorm2d (eval) is
    folded into the conv's weight and bias; LayerNorm2D is the library's LayerNorm over a frame's F*C values; the residual add
    and the un-fused ReLU are elementwise kernels."""

    def __init__(self, input_dim, in_channel, out_channel, kernel_size, stride, pooling, dropout, normalization,
                 residual):
        super().__init__()
        if tuple(kernel_size) != (3, 3):
            raise NotImplementedError("B200 front-end supports 3x3 kernels (all reference recipes)")
        if normalization not in ('', 'batch_norm', 'layer_norm'):
            raise NotImplementedError("conv_normalization=%r" % normalization)
        self.residual = residual
        self.dropout = nn.Dropout(p=dropout)
        self.time_axis = 0
        self.stride = tuple(stride)
        self.conv1 = nn.Conv2d(in_channel, out_channel, kernel_size=tuple(kernel_size), stride=(1, 1), padding=(1, 1))
        self._odim = input_dim
        self.norm1 = self._make_norm(normalization, out_channel, self._odim)
        self.conv2 = nn.Conv2d(out_channel, out_channel, kernel_size=tuple(kernel_size), stride=tuple(stride), padding=(1, 1))
        self._odim = _conv_len(self._odim, self.stride[1])
        self.norm2 = self._make_norm(normalization, out_channel, self._odim)
        self.pooling = [1, 1]
        self._factor = 1
        if len(pooling) > 0 and np.prod(pooling) > 1:
            self.pooling = list(pooling)
            self._odim = _pool_len(self._odim, pooling[1])
            if self._odim % 2 != 0:
                self._odim = (self._odim // 2) * 2
            self._factor *= pooling[0]
        self.pool = self.pooling if self._factor > 1 or self.pooling[1] > 1 else None
        # the recipes' shape: fused kernels + training path
        # the skip connection is live when the block keeps the activation's shape (conv.py:379: `xs.size() == residual.size()`)
        self.residual_active = bool(residual) and in_channel == out_channel and self.stride == (1, 1)
        # (normalised / residual blocks train too: autograd._FrontendFn runs them in fp32 with the LayerNorm / BatchNorm kernels;
        #  a skip connection around the FIRST block would add the raw feature planes: not built, see _forward_general)
        self.trainable = True
        self.plain = self.norm1 is None and not (residual and in_channel == out_channel) and self.stride == (1, 1)

    @staticmethod
    def _make_norm(normalization, channel, idim):
        if normalization == 'batch_norm':
            return nn.BatchNorm2d(channel)
        if normalization == 'layer_norm':
            return LayerNorm2D(channel, idim, eps=1e-12)
        return None

    def forward(self, xs, xlens, B, T, F, first, last_chmajor, lookback=False, lookahead=False):
        """xs: raw features (first block) or channels-last `[B, T, F, C]`.  Returns (xs, xlens, T', F').

        lookback / lookahead (streaming, reference conv.py:368-374, :385-390): after each of the two convolutions the
        leftmost / rightmost `stride` output frames -- the ones computed from zero padding instead of real neighbouring
        context -- are dropped, so a chunk fed with `context_size` extra frames on a side reproduces the offline
        activations exactly."""
        if self.training and self.dropout.p > 0:
            raise NotImplementedError("dropout > 0 in the CNN front-end is not on the B200 path (build_encoder passes 0)")
        if not self.plain:
            return self._forward_general(xs, xlens, B, T, F, first, last_chmajor, lookback, lookahead)
        prec = get_precision(self)
        adt = act_dtype(prec)
        pt, pf = self.pooling if self.pool is not None else (1, 1)
        trim = lookback or lookahead

        def conv(layer, name, x, first_layer, fuse_pool):
            ci, co = layer.in_channels, layer.out_channels
            if prec == "bf16" and ci == 32 and co == 32 and x.dtype == torch.bfloat16:
                wt = prepared(self, name + ".taps", "bf16", (layer.weight,),
                              build=lambda w: w.permute(0, 2, 3, 1).reshape(32, 288))[0][:, :288].contiguous()
                return ops.conv3x3_c32_tc(x.view(B, T, F, 32), wt, layer.bias, relu=True, pool2x2=fuse_pool), fuse_pool
            return ops.conv3x3_relu(x, layer.weight, layer.bias, B, T, F, in_chmajor=first_layer, out_dtype=adt), False

        def trim_time(x, lens):
            """Drop one frame per requested side of a channels-last `[B, T, F, C]` activation (time stride is 1)."""
            nonlocal T
            x = x.view(B, T, F, -1)
            if lookback and T > 1:
                x, T, lens = x[:, 1:], T - 1, lens - 1
            if lookahead and T > 1:
                x, T, lens = x[:, :T - 1], T - 1, lens - 1
            return x.contiguous(), lens

        xs, _ = conv(self.conv1, "conv1", xs, first, False)
        xlens = torch.IntTensor([_conv_len(int(n), 1) for n in xlens])
        if trim:
            xs, xlens = trim_time(xs, xlens)
        xs, pooled = conv(self.conv2, "conv2", xs, False, (pt, pf) == (2, 2) and not last_chmajor and not trim)
        xlens = torch.IntTensor([_conv_len(int(n), 1) for n in xlens])
        if trim:
            xs, xlens = trim_time(xs, xlens)
        if self.pool is not None:
            if not pooled:
                xs = ops.maxpool2d(xs, pt, pf, out_chmajor=last_chmajor)
            xlens = torch.IntTensor([_pool_len(int(n), pt) for n in xlens])
            T, F = -(-T // pt), -(-F // pf)
        elif last_chmajor:
            xs = ops.maxpool2d(xs, 1, 1, out_chmajor=True)
        return xs, xlens, T, F

    def _folded(self, name, conv, norm):
        """(weight, bias) of `conv`, with an eval-mode BatchNorm2d folded in: w' = w g / sigma, b' = (b - mu) g / sigma + beta."""
        if not isinstance(norm, nn.BatchNorm2d):
            return conv.weight, conv.bias
        if self.training:
            raise NotImplementedError("BatchNorm statistics updates (training mode) are not on the B200 path")

        def build(w, b, g, beta, mu, var):
            sc = g / torch.sqrt(var + norm.eps)
            return (w * sc.view(-1, 1, 1, 1)).contiguous(), ((b - mu) * sc + beta).contiguous()
        return cached(self, name + ".bnfold", (conv.weight, conv.bias, norm.weight, norm.bias, norm.running_mean,
                                               norm.running_var), build)

    def _forward_general(self, xs, xlens, B, T, F, first, last_chmajor, lookback, lookahead):
        """Strided / normalised / residual blocks (inference), fp32 channels-last activations."""
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError("this block shape trains through autograd.frontend_forward only (no normalisation, "
                                      "no residual)")
        pt, pf = self.pooling if self.pool is not None else (1, 1)
        res = None
        if self.residual and first and self.conv1.in_channels == self.conv2.out_channels:
            raise NotImplementedError("residual connection around the first CNN block (raw feature layout)")
        if self.residual and not first and self.conv1.in_channels == self.conv2.out_channels and self.stride == (1, 1):
            res = xs.reshape(B, T, F, -1).float().contiguous()       # same shape as the block's conv output (conv.py:379)

        def stage(name, conv, norm, x, first_layer, stride, residual_t, lens):
            nonlocal T, F
            w, b = self._folded(name, conv, norm)
            ln = isinstance(norm, LayerNorm2D)
            fuse_relu = not ln and residual_t is None
            y = ops.conv3x3_relu(x, w, b, B, T, F, in_chmajor=first_layer, relu=fuse_relu, out_dtype=torch.float32)
            lens = torch.IntTensor([_conv_len(int(n), stride[0]) for n in lens])
            if stride != (1, 1):                                     # 'same' 3x3 conv with stride s = stride-1 conv at 0, s, 2s, ...
                y = y[:, ::stride[0], ::stride[1]].contiguous()
                T, F = y.size(1), y.size(2)
            for flag, side in ((lookback, 0), (lookahead, 1)):       # streaming context trimming (conv.py:368-374, :385-390)
                if flag and T > stride[0]:
                    y = (y[:, stride[0]:] if side == 0 else y[:, :T - stride[0]]).contiguous()
                    T, lens = T - stride[0], lens - stride[0]
            if ln:
                gam = cached(self, name + ".ln_w", (norm.norm.weight,), lambda t: t.t().contiguous().reshape(-1).float())
                bet = cached(self, name + ".ln_b", (norm.norm.bias,), lambda t: t.t().contiguous().reshape(-1).float())
                y = ops.layernorm(y.view(B * T, -1), gam, bet, norm.norm.eps).view(B, T, F, -1)
            if residual_t is not None and residual_t.shape == y.shape:
                y = ops.dropout_add(y, residual_t, 0.0, 1.0, 0)      # p = 0: plain out = residual + y
            if not fuse_relu:
                y = ops.relu_mask(y, y)                              # relu(y) = (y > 0 ? y : 0)
            return y, lens

        xs, xlens = stage("conv1", self.conv1, self.norm1, xs if first else xs.reshape(B, T, F, -1).float().contiguous(),
                          first, (1, 1), None, xlens)
        xs, xlens = stage("conv2", self.conv2, self.norm2, xs, False, self.stride, res, xlens)
        if self.pool is not None:
            xs = ops.maxpool2d(xs, pt, pf, out_chmajor=last_chmajor)
            xlens = torch.IntTensor([_pool_len(int(n), pt) for n in xlens])
            T, F = -(-T // pt), -(-F // pf)
        elif last_chmajor:
            xs = ops.maxpool2d(xs, 1, 1, out_chmajor=True)
        return xs, xlens, T, F


def _conv1d_len(n, k, stride):      # nn.Conv1d, padding 1 (conv.py:446-450)
    return (n + 2 - (k - 1) - 1) // stride + 1


class Conv1dBlock(EncoderBase):
    """1-D CNN block over time with the feature vector as channels (reference conv.py:197-287): Conv1d(k, pad 1) -> [LayerNorm]
    -> ReLU -> Conv1d(k, stride, pad 1) -> [LayerNorm] -> [+ residual] -> ReLU -> MaxPool1d(ceil).  Each convolution is a
    tcgen05 GEMM over the k gathered neighbour frames (column order c * k + j = nn.Conv1d's weight layout), bias / ReLU in the
    epilogue.  Inference only.  `normalization='batch_norm'` is rejected: the reference applies BatchNorm1d(out_channel) to a
    `[B, T, C]` tensor, i.e. over the TIME axis (conv.py:264-266), which only runs when T == out_channel."""

    def __init__(self, in_channel, out_channel, kernel_size, stride, pooling, dropout, normalization, residual):
        super().__init__()
        if normalization == 'batch_norm':
            raise NotImplementedError("conv_normalization=batch_norm with a 1-D CNN front-end (see the class docstring)")
        self.residual = residual
        self.dropout = nn.Dropout(p=dropout)
        self.kernel_size, self.stride, self.pooling = kernel_size, stride, pooling
        self.conv1 = nn.Conv1d(in_channel, out_channel, kernel_size=kernel_size, stride=1, padding=1)
        self.norm1 = nn.LayerNorm(out_channel, eps=1e-12) if normalization == 'layer_norm' else None
        self.conv2 = nn.Conv1d(out_channel, out_channel, kernel_size=kernel_size, stride=stride, padding=1)
        self.norm2 = nn.LayerNorm(out_channel, eps=1e-12) if normalization == 'layer_norm' else None
        self.pool = pooling if pooling > 1 else None
        self._odim = out_channel
        self.plain = self.trainable = False     # trains through train_forward, not through the fused 2-D node

    def _conv(self, name, conv, norm, xs, stride, residual_t):
        prec = get_precision(self)
        B, T, C = xs.shape
        k = self.kernel_size
        To = _conv1d_len(T, k, stride)
        xp = torch.nn.functional.pad(xs, (0, 0, 1, 1))                         # layout plumbing only
        cols = xp.unfold(1, k, stride)[:, :To].reshape(B, To, C * k)           # column c * k + j
        w = prepared(self, name, prec, (conv.weight,), build=lambda t: t.reshape(t.size(0), -1))
        fuse_relu = norm is None and residual_t is None
        y = ops.linear(cols, w, conv.bias, prec=prec, act="relu" if fuse_relu else None, out_dtype=torch.float32)
        if norm is not None:
            y = ops.layernorm(y, norm.weight, norm.bias, norm.eps)
        if residual_t is not None and residual_t.shape == y.shape:
            y = ops.dropout_add(y.contiguous(), residual_t, 0.0, 1.0, 0)       # p = 0: plain out = residual + y
        if not fuse_relu:
            y = ops.relu_mask(y.contiguous(), y.contiguous())
        return y

    def train_forward(self, xs, xlens):
        """Training path (autograd nodes: GEMM + fused ReLU with tcgen05 dgrad / wgrad, time max-pool); the un-normalised,
        non-residual block shape only."""
        from .. import autograd as ag
        if self.norm1 is not None or self.residual:
            raise NotImplementedError("training path of the 1-D CNN front-end: no normalisation, no residual only")
        if self.dropout.p > 0:
            raise NotImplementedError("dropout > 0 in the 1-D CNN front-end is not on the B200 training path "
                                      "(build_encoder passes 0; the reference applies it after each ReLU, conv.py:265-283)")
        prec = get_precision(self)
        k = self.kernel_size
        for name, conv, stride in (("conv1", self.conv1, 1), ("conv2", self.conv2, self.stride)):
            B, T, C = xs.shape
            To = _conv1d_len(T, k, stride)
            cols = torch.nn.functional.pad(xs, (0, 0, 1, 1)).unfold(1, k, stride)[:, :To].reshape(B, To, C * k)
            xs = ag.linear_relu(self, name + "_ck", conv.weight, conv.bias, cols, prec)
            xlens = torch.IntTensor([_conv1d_len(int(n), k, stride) for n in xlens])
        if self.pool is not None:
            xs = ag.maxpool_time(xs, self.pool)
            xlens = torch.IntTensor([_pool_len(int(n), self.pool) for n in xlens])
        return xs, xlens

    def forward(self, xs, xlens, lookback=False, lookahead=False):
        """xs fp32 `[B, T, C_in]` -> (`[B, T', C_out]`, xlens); lookback / lookahead are accepted and unused, as in the reference."""
        if self.training and torch.is_grad_enabled():
            return self.train_forward(xs.float(), xlens)
        k = self.kernel_size
        res = xs.float().contiguous() if self.residual else None
        xs = self._conv("conv1", self.conv1, self.norm1, xs.float(), 1, None)
        xlens = torch.IntTensor([_conv1d_len(int(n), k, 1) for n in xlens])
        xs = self._conv("conv2", self.conv2, self.norm2, xs, self.stride, res)
        xlens = torch.IntTensor([_conv1d_len(int(n), k, self.stride) for n in xlens])
        if self.pool is not None:
            xs = ops.pool_time(xs.contiguous(), self.pool, "max")
            xlens = torch.IntTensor([_pool_len(int(n), self.pool) for n in xlens])
        return xs, xlens


class ConvEncoder(EncoderBase):
    def __init__(self, input_dim, in_channel, channels, kernel_sizes, strides, poolings, dropout, normalization,
                 residual, bottleneck_dim, param_init):
        super().__init__()
        assert channels
        (channels, kernel_sizes, strides, poolings), is_1dconv = parse_cnn_config(channels, kernel_sizes, strides, poolings)
        self.is_1dconv = is_1dconv
        self.in_channel = in_channel
        assert input_dim % in_channel == 0
        self.input_freq = input_dim // in_channel
        self.residual = residual
        assert len(channels) > 0 and len(channels) == len(kernel_sizes) == len(strides) == len(poolings)
        self.layers = nn.ModuleList()
        C_i, in_freq = (input_dim if is_1dconv else in_channel), self.input_freq
        for lth in range(len(channels)):
            if is_1dconv:               # features as channels, convolution over time only (reference :60-68)
                block = Conv1dBlock(C_i, channels[lth], kernel_sizes[lth], strides[lth], poolings[lth], dropout,
                                    normalization, residual)
            else:
                block = Conv2dBlock(in_freq, C_i, channels[lth], kernel_sizes[lth], strides[lth], poolings[lth],
                                    dropout, normalization, residual)
            self.layers += [block]
            in_freq = block.output_dim
            C_i = channels[lth]
        self._c_last, self._f_last = C_i, in_freq
        self._odim = C_i if is_1dconv else int(C_i * in_freq)
        self.bridge = None
        if bottleneck_dim > 0 and bottleneck_dim != self._odim:
            self.bridge = nn.Linear(self._odim, bottleneck_dim)
            self._odim = bottleneck_dim
        self._factor = 1
        for s_, p_ in zip(strides, poolings):
            self._factor *= (s_ if is_1dconv else s_[0]) * (p_ if is_1dconv else p_[0])
        if is_1dconv:
            kernel_sizes, strides, poolings = ([[v] for v in lst] for lst in (kernel_sizes, strides, poolings))
        self._context_size = self._calc_context(kernel_sizes, strides, poolings)
        for n, p in self.named_parameters():
            init_with_lecun_normal(n, p, param_init)

    @staticmethod
    def _calc_context(kernel_sizes, strides, poolings):     # reference conv.py:142-159
        ctx, bottom, factor = 0, 0, 1
        for ks, st, po in zip(kernel_sizes, strides, poolings):
            look = ((ks[0] - 1) // 2) * 2
            if factor == 1:
                ctx += look
                bottom = ctx
            else:
                ctx += bottom * look
                bottom *= st[0] * po[0]
            factor *= st[0] * po[0]
        return ctx

    @property
    def context_size(self):
        return self._context_size

    def output_lens(self, xlens):
        """Length arithmetic of the block stack alone (reference conv.py:451-477), for the training path."""
        for block in self.layers:
            if not block.trainable:
                raise NotImplementedError("training path of the CNN front-end: no normalisation, no residual only")
            xlens = torch.IntTensor([_conv_len(_conv_len(int(n), 1), block.stride[0]) for n in xlens])
            if block.pool is not None:
                xlens = torch.IntTensor([_pool_len(int(n), block.pooling[0]) for n in xlens])
        return xlens

    def forward(self, xs, xlens, lookback=False, lookahead=False, out_scale=1.0):
        """xs `[B, T, F]` fp32 on the GPU, xlens IntTensor `[B]` (CPU) -> (`[B, T', odim]` fp32, xlens)."""
        B, T, Fdim = xs.size()
        F = Fdim // self.in_channel
        prec = get_precision(self)
        xs = xs.contiguous().float()
        n = len(self.layers)
        if self.is_1dconv:
            if self.training and torch.is_grad_enabled():
                from .. import autograd as ag
                for block in self.layers:
                    xs, xlens = block.train_forward(xs, xlens)
                if self.bridge is not None:
                    xs = ag.linear(self, "bridge1d", self.bridge, xs, prec)
                return ag.scale(xs, out_scale), xlens
            for block in self.layers:
                xs, xlens = block(xs, xlens, lookback=lookback, lookahead=lookahead)
            if self.bridge is not None:
                xs = ops.linear(xs, prepared(self, "bridge1d", prec, (self.bridge.weight,)), self.bridge.bias, prec=prec,
                                alpha=out_scale, out_dtype=torch.float32)
            elif out_scale != 1.0:
                xs = ops.scale_(xs.contiguous(), out_scale)
            return xs, xlens
        for i, block in enumerate(self.layers):
            # with a bridge the channels-last flatten is absorbed by permuting the bridge weight's columns
            last_chmajor = (i == n - 1) and self.bridge is None
            xs, xlens, T, F = block(xs, xlens, B, T, F, first=(i == 0), last_chmajor=last_chmajor,
                                    lookback=lookback, lookahead=lookahead)
        if self.bridge is not None:
            C, Fo = self._c_last, F
            wb = prepared(self, "bridge", prec, (self.bridge.weight,),
                          build=lambda w: w.view(w.size(0), C, Fo).transpose(1, 2).reshape(w.size(0), Fo * C))
            xs = ops.linear(xs.reshape(B, T, Fo * C), wb, self.bridge.bias, prec=prec, alpha=out_scale,
                            out_dtype=torch.float32)
        else:
            xs = xs.float() if xs.dtype != torch.float32 else xs
            if out_scale != 1.0:
                xs = ops.scale_(xs.contiguous(), out_scale)
        return xs, xlens


# command-line contract of the reference (add_args / define_name static methods): see encoders/cli.py
from . import cli as _cli  # noqa: E402

ConvEncoder.add_args = staticmethod(_cli.conv_add_args)
ConvEncoder.define_name = staticmethod(_cli.conv_define_name)
