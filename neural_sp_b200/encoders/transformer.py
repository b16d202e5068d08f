"""Transformer encoder (reference encoders/transformer.py:40-617), B200-native.

Same constructor arguments, ``forward(xs, xlens, task, ...) -> {'ys': {'xs', 'xlens'}, 'ys_sub1', 'ys_sub2'}``
and state_dict keys (``conv.*``, ``embed.*``, ``pos_emb.inv_freq``, ``u_bias/v_bias``, ``layers.N.*``,
``norm_out.*``, ``bridge*.*``).  The `[B,T',T']` boolean mask of the reference (:633-686) is never built:
key-padding / causal / chunk visibility are evaluated inside the attention kernel from device-side lengths.
Supported here: offline full-context, unidirectional ('uni', per-layer lookahead) and latency-controlled
('reshape' overlapped windows / 'mask' chunk-wise visibility) encoders with all six hierarchical subsamplers and
sub-task outputs, and chunk-by-chunk streaming inference (``streaming=True``, :448-606): per-layer caches of the
normalised attention input (and of the Conformer conv-module input) are carried in ``self.cache`` exactly like the
reference's -- same keys, same truncation to ``cache_sizes``; one private extra entry ``_kv`` holds the already PROJECTED
keys / values of the cached frames, so a chunk costs O(chunk) GEMM work instead of re-projecting the whole cache as the
reference does -- and the attention kernel sees the cached frames as
`mlen = klen - qlen` extra keys (query i sits at key position mlen + i for the causal / chunk masks and the relative
distance), so nothing of the `[B, qlen, klen]` mask or the shifted position matrix is ever materialised."""
import copy
import math

import numpy as np
import torch
import torch.nn as nn

import random

from .. import autograd as ag
from .. import ops
from ..modules._prep import prepared, get_precision
from ..modules.positional_embedding import PositionalEncoding, XLPositionalEmbedding
from .encoder_base import EncoderBase
from .subsampling import (AddSubsampler, ConcatSubsampler, Conv1dSubsampler, DropSubsampler, MaxPoolSubsampler,
                          MeanPoolSubsampler)
from .transformer_block import TransformerEncoderBlock


_LENS_CACHE = {}


def chunkwise(xs, N_l, N_c, N_r):
    """`[B, T, D]` -> `[B * ceil(T / N_c), N_l + N_c + N_r, D]` overlapped windows, zero padded at both ends
    (reference encoders/utils.py:13-45, padding=True).  Pure data movement."""
    bs, xmax, idim = xs.size()
    n_chunks = math.ceil(xmax / N_c)
    width = N_l + N_c + N_r
    xp = torch.nn.functional.pad(xs, (0, 0, N_l, n_chunks * N_c - xmax + N_r))
    out = xp.unfold(1, width, N_c)[:, :n_chunks]                 # [B, n_chunks, D, width]
    return out.permute(0, 1, 3, 2).reshape(bs * n_chunks, width, idim)


def lens_to_device(xlens, device):
    """CPU IntTensor -> int32 CUDA tensor without a host sync.  Results are cached per (lengths, device): repeated
    batches of the same lengths (fixed-shape benchmarking, CUDA-graph replays) reuse the device copy."""
    key = (tuple(int(v) for v in xlens), str(device))
    hit = _LENS_CACHE.get(key)
    if hit is not None:
        return hit
    t = xlens.to(torch.int32)
    if device.type == "cuda":
        t = t.pin_memory()
    d = t.to(device, non_blocking=True)
    if len(_LENS_CACHE) > 4096:
        _LENS_CACHE.clear()
    _LENS_CACHE[key] = d
    return d


class TransformerEncoder(EncoderBase):
    def __init__(self, input_dim, enc_type, n_heads, n_layers, n_layers_sub1, n_layers_sub2, d_model, d_ff,
                 ffn_bottleneck_dim, ffn_activation, pe_type, layer_norm_eps, last_proj_dim, dropout_in, dropout,
                 dropout_att, dropout_layer, subsample, subsample_type, n_stacks, n_splices, frontend_conv,
                 task_specific_layer, param_init, clamp_len, lookahead, chunk_size_left, chunk_size_current,
                 chunk_size_right, streaming_type):
        super().__init__()
        self.subsample_factors = [1] * n_layers
        for lth, s in enumerate(list(map(int, subsample.split('_')[:n_layers]))):
            self.subsample_factors[lth] = s
        lookaheads = [0] * n_layers
        for lth, s in enumerate(list(map(int, lookahead.split('_')[:n_layers]))):
            lookaheads[lth] = s
        self.enc_type = enc_type
        self.d_model = d_model
        self.n_layers = n_layers
        self.n_heads = n_heads
        self.pe_type = pe_type
        self.scale = math.sqrt(d_model)
        self.unidir = 'uni' in enc_type
        self.lookaheads = lookaheads
        if sum(lookaheads) > 0:
            assert self.unidir
        cl, cc, cr = str(chunk_size_left), str(chunk_size_current), str(chunk_size_right)
        self.N_l = int(cl.split('_')[-1]) // n_stacks
        self.N_c = int(cc.split('_')[-1]) // n_stacks
        self.N_r = int(cr.split('_')[-1]) // n_stacks
        self.lc_bidir = self.N_c > 0 and enc_type != 'conv' and 'uni' not in enc_type
        self.cnn_lookahead = self.unidir or enc_type == 'conv'
        self.streaming_type = streaming_type if self.lc_bidir else ''
        self.causal = self.unidir or self.streaming_type == 'mask'
        if self.lc_bidir:
            assert n_layers_sub1 == 0 and n_layers_sub2 == 0 and not self.unidir
        if self.streaming_type == 'mask':
            assert self.N_r == 0 and self.N_l % self.N_c == 0
        self.n_layers_sub1 = n_layers_sub1
        self.n_layers_sub2 = n_layers_sub2
        self.task_specific_layer = task_specific_layer
        self.bridge = self.bridge_sub1 = self.bridge_sub2 = None
        self.aws_dict, self.data_dict = {}, {}

        self.conv = frontend_conv
        if self.conv is not None:
            self._odim = self.conv.output_dim
        else:
            self._odim = input_dim * n_splices * n_stacks
            self.embed = nn.Linear(self._odim, d_model)
        self._factor = 1
        self.conv_factor = self.conv.subsampling_factor if self.conv is not None else 1
        self._factor *= self.conv_factor
        self.subsample_layers = None
        if np.prod(self.subsample_factors) > 1:
            self._factor *= int(np.prod(self.subsample_factors))
            odim = self._odim
            make = {'max_pool': MaxPoolSubsampler, 'mean_pool': MeanPoolSubsampler, 'drop': DropSubsampler,
                    'add': AddSubsampler, 'concat': lambda f: ConcatSubsampler(f, odim),
                    'conv1d': lambda f: Conv1dSubsampler(f, odim)}
            if subsample_type not in make:
                raise NotImplementedError(subsample_type)
            if subsample_type == 'conv1d':
                assert not self.causal
            self.subsample_layers = nn.ModuleList([make[subsample_type](f) for f in self.subsample_factors])

        self.pos_enc, self.pos_emb = None, None
        self.u_bias, self.v_bias = None, None
        if pe_type in ['relative', 'relative_xl']:
            self.pos_emb = XLPositionalEmbedding(d_model, dropout)
            if pe_type == 'relative_xl':
                self.u_bias = nn.Parameter(torch.Tensor(n_heads, d_model // n_heads))
                self.v_bias = nn.Parameter(torch.Tensor(n_heads, d_model // n_heads))
        else:
            self.pos_enc = PositionalEncoding(d_model, dropout_in, pe_type, param_init)

        self.layers = nn.ModuleList([copy.deepcopy(TransformerEncoderBlock(
            d_model, d_ff, n_heads, dropout, dropout_att, dropout_layer * (lth + 1) / n_layers, layer_norm_eps,
            ffn_activation, param_init, pe_type, clamp_len, ffn_bottleneck_dim)) for lth in range(n_layers)])
        self.norm_out = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self._odim = d_model

        for sub, nl in (('sub1', n_layers_sub1), ('sub2', n_layers_sub2)):
            if nl > 0:
                if task_specific_layer:         # one extra block on top of layer `nl` for the sub task (reference :232-238)
                    setattr(self, 'layer_' + sub, TransformerEncoderBlock(
                        d_model, d_ff, n_heads, dropout, dropout_att, dropout_layer * nl / n_layers, layer_norm_eps,
                        ffn_activation, param_init, pe_type, clamp_len, ffn_bottleneck_dim))
                odim_sub = d_model
                if last_proj_dim > 0 and last_proj_dim != self.output_dim:
                    setattr(self, 'bridge_' + sub, nn.Linear(self._odim, last_proj_dim))
                    odim_sub = last_proj_dim
                setattr(self, 'norm_out_' + sub, None if nl == n_layers else nn.LayerNorm(odim_sub, eps=layer_norm_eps))
        if last_proj_dim > 0 and last_proj_dim != self.output_dim:
            self.bridge = nn.Linear(self._odim, last_proj_dim)
            self._odim = last_proj_dim
        self.reset_parameters(param_init)
        self.reset_cache()
        self.cache_sizes = self.calculate_cache_size()

    def reset_parameters(self, param_init):
        if param_init == 'xavier_uniform':
            lins = [getattr(self, 'embed', None), self.bridge, self.bridge_sub1, self.bridge_sub2]
            for lin in lins:
                if lin is not None:
                    nn.init.xavier_uniform_(lin.weight)
                    nn.init.constant_(lin.bias, 0.)
            if self.pe_type == 'relative_xl':
                nn.init.xavier_uniform_(self.u_bias)
                nn.init.xavier_uniform_(self.v_bias)

    def reset_cache(self):
        """Reset the streaming state (reference transformer.py:370-374)."""
        self.cache = [None] * self.n_layers
        self.offset = 0

    def truncate_cache(self, cache):
        """Keep at most ``cache_sizes[lth]`` cached attention-input frames per layer (reference :376-391)."""
        if cache[0] is not None:
            for lth in range(self.n_layers):
                size = self.cache_sizes[lth]
                san = cache[lth]['input_san']
                if san.size(1) > size:
                    cache[lth]['input_san'] = san[:, san.size(1) - size:]
                    if '_kv' in cache[lth]:     # the projected keys / values of the same frames (kept next to input_san)
                        cache[lth]['_kv'] = tuple(t[:, t.size(1) - size:] for t in cache[lth]['_kv'])
        return cache

    def calculate_cache_size(self):
        """Maximum number of cached frames per layer, after CNN subsampling (reference :393-404)."""
        size = self._total_chunk_size_left()
        N_l = self.N_l // self.conv_factor
        sizes = []
        for lth in range(self.n_layers):
            sizes.append(size)
            if self.lc_bidir:
                size = max(0, size - N_l)
                N_l //= self.subsample_factors[lth]
            size //= self.subsample_factors[lth]
        return sizes

    def _total_chunk_size_left(self):
        """Left context accumulated over the layer stack, in frames after the CNN (reference :406-417)."""
        if self.streaming_type == 'reshape':
            return self.N_l // self.conv_factor
        if self.streaming_type == 'mask':
            return (self.N_l // self.conv_factor) * self.n_layers
        return 10000 // self.conv_factor

    def _proj(self, name, lin, xs, scale=1.0, train=False):
        prec = get_precision(self)
        if train:
            return ag.scale(ag.linear(self, name, lin, xs, prec), scale)
        return ops.linear(xs, prepared(self, name, prec, (lin.weight,)), lin.bias, prec=prec, alpha=scale,
                          out_dtype=torch.float32)

    def _norm(self, norm, xs, train=False):
        if train:
            return ag.layernorm(norm, xs)
        return ops.layernorm(xs, norm.weight, norm.bias, norm.eps)

    def _sub_out(self, xs, module, train=False, klens=None, pos=None, mask_kw=None):
        if self.task_specific_layer:            # reference sub_module (:619-625): the extra block sees no u / v bias
            layer = getattr(self, 'layer_' + module)
            if train:
                xs_sub = self._train_layer(-1, layer, xs, klens, pos, mask_kw, get_precision(self), rel_bias=(None, None))
            else:
                xs_sub, _ = layer(xs.clone(), klens, cache=None, pos_embs=pos, rel_bias=(None, None), mask_kw=mask_kw)
        else:
            xs_sub = xs.clone()
        bridge = getattr(self, 'bridge_' + module)
        if bridge is not None:
            xs_sub = self._proj('bridge_' + module, bridge, xs_sub, train=train)
        norm = getattr(self, 'norm_out_' + module)
        if norm is not None:
            xs_sub = self._norm(norm, xs_sub, train)
        return xs_sub

    def _train_layer(self, lth, layer, xs, klens, pos, mask_kw, prec, rel_bias=None):
        """One block through its autograd node (training): LayerDrop as in conformer_block.py:122-126."""
        if layer.self_attn.dropout_attn.p > 0:
            raise NotImplementedError("dropout on the attention weights (dropout_att > 0) is not on the B200 path: the "
                                      "flash-style kernels never materialise them (the LibriSpeech recipes use 0)")
        in_scale = 1.0
        if layer.dropout_layer > 0:
            if random.random() < layer.dropout_layer:
                return xs
            in_scale = 1.0 / (1 - layer.dropout_layer)
        rb = (self.u_bias, self.v_bias) if rel_bias is None else rel_bias
        return ag.block_forward(layer, xs, klens, pos, rb, mask_kw, prec, in_scale)

    def forward(self, xs, xlens, task, streaming=False, lookback=False, lookahead=False):
        """xs `[B, T, input_dim]` fp32 on the GPU; xlens IntTensor `[B]` on the CPU (reference contract).
        streaming=True encodes one chunk and carries ``self.cache`` / ``self.offset`` over to the next call
        (``reset_cache()`` starts a new utterance); lookback / lookahead trim the CNN context frames."""
        eouts = {'ys': {'xs': None, 'xlens': None}, 'ys_sub1': {'xs': None, 'xlens': None},
                 'ys_sub2': {'xs': None, 'xlens': None}}
        # train() + grad mode: every block / the front-end is one autograd node with a hand-written CUDA backward
        # (neural_sp_b200/autograd.py); otherwise inference kernels under no_grad (in-place residual stream).
        train = ag.training_enabled(self)
        if streaming and train:
            raise NotImplementedError("streaming=True is an inference path (call .eval() / torch.no_grad())")
        prec = get_precision(self)
        with (torch.enable_grad() if train else torch.no_grad()):
            rel = 'relative' in self.pe_type
            bs, xmax = xs.size(0), xs.size(1)
            N_l, N_c, N_r = self.N_l, self.N_c, self.N_r
            st = self.streaming_type
            if streaming and st == 'mask':
                assert xmax <= N_c
            elif streaming and st == 'reshape':
                assert xmax <= N_l + N_c + N_r
            if self.lc_bidir:                                         # latency-controlled encoders (:456-465)
                if st == 'mask' and not streaming:
                    xs = chunkwise(xs, 0, N_c, 0)
                elif st == 'reshape' and not streaming:
                    xs = chunkwise(xs, N_l, N_c, N_r)
                elif st == 'reshape':                                 # one window, already carrying both contexts
                    assert xmax // (N_l + N_c + N_r) == 1, xs.size()
                    xs = xs[:, :N_l + N_c + N_r]
            if self.conv is None:
                xs = self._proj('embed', self.embed, xs.float(), scale=self.scale if rel else 1.0, train=train)
            elif train:
                if lookback or lookahead:
                    raise NotImplementedError("CNN lookback/lookahead trimming is an inference (streaming) feature")
                if self.conv.is_1dconv:       # composed of GEMM / pooling nodes inside the front-end module
                    xs, xlens = self.conv(xs, xlens, out_scale=self.scale if (rel and self.enc_type != 'conv') else 1.0)
                else:
                    xs = ag.frontend_forward(self.conv, xs, self.scale if (rel and self.enc_type != 'conv') else 1.0, prec)
                    xlens = self.conv.output_lens(xlens)
                N_l, N_c, N_r = max(0, N_l // self.conv_factor), N_c // self.conv_factor, N_r // self.conv_factor
            else:
                xs, xlens = self.conv(xs, xlens, lookback=False if self.lc_bidir else lookback,
                                      lookahead=False if self.lc_bidir else lookahead,
                                      out_scale=self.scale if (rel and self.enc_type != 'conv') else 1.0)
                N_l, N_c, N_r = max(0, N_l // self.conv_factor), N_c // self.conv_factor, N_r // self.conv_factor
            emax = xs.size(1)
            if streaming and st != 'reshape':                         # (:476-478)
                xlens = torch.IntTensor([int(v) for v in xlens])
                xs = xs[:, :int(xlens.max())].contiguous()
                xlens = xlens.clamp(max=xs.size(1))
            elif not streaming and st == 'mask':                      # back to utterance shape (:479-481)
                xs = xs.contiguous().view(bs, -1, xs.size(2))[:, :int(xlens.max())].contiguous()
            if self.enc_type == 'conv':
                eouts['ys']['xs'], eouts['ys']['xlens'] = xs, xlens
                return eouts
            if streaming:
                self.cache = self.truncate_cache(self.cache)
            else:
                self.reset_cache()
            dev = xs.device
            if not rel:                                               # absolute positions: xs * sqrt(d) [+ pe] (:497-498)
                if train:
                    xs = ag.dropout(ag.add_pos_enc(self.pos_enc, xs.contiguous(), self.offset), self.pos_enc.dropout.p)
                else:
                    xs = self.pos_enc(xs.contiguous(), scale=True, offset=self.offset)

            def cached_frames(lth):
                c = self.cache[lth] if streaming else None
                return c['input_san'].size(1) if c is not None else 0

            def key_lens(n_cache):
                if st == 'reshape':                                   # no mask at all inside a chunk (:512)
                    return lens_to_device(torch.IntTensor([xs.size(1)] * xs.size(0)), dev)
                return lens_to_device(xlens + n_cache if n_cache else xlens, dev)

            def pos_table(rows):
                """Sinusoid rows 0..rows-1; training: dropped like the reference's `self.dropout(pos_emb)`
                (positional_embedding.py:139) -- a constant, so no gradient flows into it."""
                tab = self.pos_emb.table(rows)
                if train and self.pos_emb.dropout.p > 0:
                    from .. import random as nrandom
                    tab = ops.dropout(tab.contiguous(), self.pos_emb.dropout.p, nrandom.next_stream())
                return tab

            n_cache = cached_frames(0)
            klens = key_lens(n_cache)
            pos = pos_table(xs.size(1) + n_cache) if rel else None
            new_cache = [None] * self.n_layers

            def mask_kw(lth):
                if self.unidir:
                    return dict(causal=True, lookahead=self.lookaheads[lth])
                if st == 'mask':
                    return dict(chunk_c=N_c, chunk_l=N_l)
                return {}

            for lth, layer in enumerate(self.layers):
                if train:
                    xs = self._train_layer(lth, layer, xs, klens, pos, mask_kw(lth), prec)
                else:
                    xs, cache = layer(xs, klens, cache=self.cache[lth] if streaming else None, pos_embs=pos,
                                      rel_bias=(self.u_bias, self.v_bias), mask_kw=mask_kw(lth))
                    if streaming and st != 'reshape':                 # 'reshape' windows carry their own context
                        new_cache[lth] = cache
                if lth == self.n_layers_sub1 - 1:
                    xs_sub1, xlens_sub1 = self._sub_out(xs, 'sub1', train, klens, pos, mask_kw(lth)), xlens.clone()
                    if task == 'ys_sub1':
                        eouts[task]['xs'], eouts[task]['xlens'] = xs_sub1, xlens_sub1
                        return eouts
                if lth == self.n_layers_sub2 - 1:
                    xs_sub2, xlens_sub2 = self._sub_out(xs, 'sub2', train, klens, pos, mask_kw(lth)), xlens.clone()
                    if task == 'ys_sub2':
                        eouts[task]['xs'], eouts[task]['xlens'] = xs_sub2, xlens_sub2
                        return eouts
                if lth < len(self.layers) - 1:
                    f = self.subsample_factors[lth]
                    if f > 1:
                        if train:
                            xs, xlens = ag.subsample_train(self.subsample_layers[lth], xs, xlens)
                        else:
                            xs, xlens = self.subsample_layers[lth](xs, xlens)
                        N_l, N_c, N_r = max(0, N_l // f), N_c // f, N_r // f
                    if streaming or f > 1:                            # cache sizes differ per layer (:535-541, :582-588)
                        n_cache = cached_frames(lth + 1)
                        klens = key_lens(n_cache)
                        if rel:
                            pos = pos_table(xs.size(1) + n_cache)
            if st == 'reshape':                                       # keep the centre of every window (:546-550)
                xs = xs[:, N_l:N_l + N_c].contiguous().view(bs, -1, xs.size(2))[:, :int(xlens.max())].contiguous()
            xs = self._norm(self.norm_out, xs, train)
            if streaming:
                self.cache = new_cache
                if st != 'reshape':
                    self.offset += emax
            if self.bridge is not None:
                xs = self._proj('bridge', self.bridge, xs, train=train)
        if task in ['all', 'ys']:
            eouts['ys']['xs'], eouts['ys']['xlens'] = xs, xlens
        if self.n_layers_sub1 >= 1 and task == 'all':
            eouts['ys_sub1']['xs'], eouts['ys_sub1']['xlens'] = xs_sub1, xlens_sub1
        if self.n_layers_sub2 >= 1 and task == 'all':
            eouts['ys_sub2']['xs'], eouts['ys_sub2']['xlens'] = xs_sub2, xlens_sub2
        return eouts


# command-line contract of the reference (add_args / define_name static methods): see encoders/cli.py
from . import cli as _cli  # noqa: E402

TransformerEncoder.add_args = staticmethod(_cli.transformer_add_args)
TransformerEncoder.define_name = staticmethod(_cli.transformer_define_name)
