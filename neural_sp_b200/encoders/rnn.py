"""(CNN-)LSTM / BLSTM encoder (reference encoders/rnn.py:35-510), B200-native.

Same constructor arguments and state_dict keys (``rnn.N.weight_ih_l0[_reverse]`` ..., ``proj.N.*``, ``subsample.N.*``,
``bridge*.*``, ``conv.*``); the nn.LSTM modules only hold the parameters.  Per layer the input projection of all frames
is one tcgen05 GEMM (both directions and both biases folded in) and the recurrence is the persistent LSTM kernel of the
library, which implements the packed-sequence semantics of ``Padding`` (:534-546) from device-side lengths -- so the
reference's sort / pack / unsort round trip (:296-298, :375-377) is unnecessary.
Supported: lstm / blstm (+ conv front-end), projections, sum of directions, the six subsamplers, sub-task outputs,
bridge.  In train() + grad mode every LSTM layer is one autograd node (neural_sp_b200/autograd.py: forward keeps the gate
activations, backward = persistent BPTT kernel + tcgen05 dgrad / wgrad GEMMs); all six subsamplers have a training path.
Streaming inference (``streaming=True``, :343-346): every layer's (h_n, c_n) is carried in ``self.hx_fwd`` across chunks by
the kernel's initial / final state arguments; like the reference's un-packed ``rnn(xs, hx)`` call, a streamed chunk is
processed over all of its frames.  Latency-controlled BLSTM (``chunk_size_current/right`` with a bidirectional type,
:427-510): separate ``rnn`` / ``rnn_bwd`` unidirectional LSTMs per layer, the forward one carrying its state over N_c
frames and looking N_r frames ahead, the backward one restarted on every chunk; offline it loops over the chunks of the
utterance, streaming it encodes one; in training the forward LSTM's state stays attached to the graph from chunk to chunk
(the BPTT kernel takes / returns state gradients).  Random state passing is not on the B200 path.  The reference sorts the batch by length before the layers and un-sorts afterwards (:296-298, :375-377);
nothing here needs the sort, so carried states stay in the caller's batch order."""
import math

import numpy as np
import torch
import torch.nn as nn

from .. import autograd as ag
from .. import ops
from ..modules._prep import prepared, cached, get_precision
from .encoder_base import EncoderBase
from .subsampling import (AddSubsampler, ConcatSubsampler, Conv1dSubsampler, DropSubsampler, MaxPoolSubsampler,
                          MeanPoolSubsampler)
from .transformer import chunkwise, lens_to_device


class RNNEncoder(EncoderBase):
    def __init__(self, input_dim, enc_type, n_units, n_projs, last_proj_dim, n_layers, n_layers_sub1, n_layers_sub2,
                 dropout_in, dropout, subsample, subsample_type, n_stacks, n_splices, frontend_conv, bidir_sum_fwd_bwd,
                 task_specific_layer, param_init, chunk_size_current, chunk_size_right, cnn_lookahead, rsp_prob):
        super().__init__()
        subsamples = [1] * n_layers
        for lth, s in enumerate(list(map(int, subsample.split('_')[:n_layers]))):
            subsamples[lth] = s
        self.enc_type = enc_type
        self.bidirectional = 'blstm' in enc_type
        self.n_units = n_units
        self.n_dirs = 2 if self.bidirectional else 1
        self.n_layers = n_layers
        self.bidir_sum = bidir_sum_fwd_bwd
        self.N_c = int(str(chunk_size_current).split('_')[0]) // n_stacks
        self.N_r = int(str(chunk_size_right).split('_')[0]) // n_stacks
        self.lc_bidir = (self.N_c > 0 or self.N_r > 0) and self.bidirectional
        if self.lc_bidir:
            assert enc_type not in ['lstm', 'conv_lstm'] and n_layers_sub2 == 0
        self.rsp_prob = rsp_prob        # random state passing acts in training only (rnn.py:323-324); see forward()
        self.n_layers_sub1, self.n_layers_sub2 = n_layers_sub1, n_layers_sub2
        self.task_specific_layer = task_specific_layer
        self.bridge = self.bridge_sub1 = self.bridge_sub2 = None
        self.dropout_in = nn.Dropout(p=dropout_in)
        self.conv = frontend_conv
        self._odim = self.conv.output_dim if self.conv is not None else input_dim * n_splices * n_stacks
        self.cnn_lookahead = cnn_lookahead
        if not cnn_lookahead:
            assert self.N_c > 0 and self.lc_bidir
        if enc_type != 'conv':
            self.rnn = nn.ModuleList()
            if self.lc_bidir:
                self.rnn_bwd = nn.ModuleList()
            self.dropout = nn.Dropout(p=dropout)
            self.proj = nn.ModuleList() if n_projs > 0 else None
            self.subsample = nn.ModuleList() if np.prod(subsamples) > 1 else None
            for lth in range(n_layers):
                if self.lc_bidir:                   # two unidirectional LSTMs per layer (reference :143-147)
                    self.rnn += [nn.LSTM(self._odim, n_units, 1, batch_first=True)]
                    self.rnn_bwd += [nn.LSTM(self._odim, n_units, 1, batch_first=True)]
                else:
                    self.rnn += [nn.LSTM(self._odim, n_units, 1, batch_first=True, bidirectional=self.bidirectional)]
                self._odim = n_units if bidir_sum_fwd_bwd else n_units * self.n_dirs
                for sub, nl in (('sub1', n_layers_sub1), ('sub2', n_layers_sub2)):
                    if lth == nl - 1 and task_specific_layer:
                        setattr(self, 'layer_' + sub, nn.Linear(self._odim, n_units))
                        setattr(self, '_odim_' + sub, n_units)
                        if last_proj_dim > 0 and last_proj_dim != self.output_dim:
                            setattr(self, 'bridge_' + sub, nn.Linear(n_units, last_proj_dim))
                            setattr(self, '_odim_' + sub, last_proj_dim)
                if self.proj is not None and lth != n_layers - 1:
                    self.proj += [nn.Linear(self._odim, n_projs)]
                    self._odim = n_projs
                if self.subsample is not None:
                    odim = self._odim
                    make = {'max_pool': MaxPoolSubsampler, 'mean_pool': MeanPoolSubsampler, 'drop': DropSubsampler,
                            'add': AddSubsampler, 'concat': lambda f: ConcatSubsampler(f, odim),
                            'conv1d': lambda f: Conv1dSubsampler(f, odim)}
                    self.subsample += [make[subsample_type](subsamples[lth])]
            if last_proj_dim > 0 and last_proj_dim != self.output_dim:
                self.bridge = nn.Linear(self._odim, last_proj_dim)
                self._odim = last_proj_dim
        self.conv_factor = self.conv.subsampling_factor if self.conv is not None else 1
        self._factor = self._factor_sub1 = self._factor_sub2 = self.conv_factor
        if n_layers_sub1 > 1:
            self._factor_sub1 *= int(np.prod(subsamples[:n_layers_sub1 - 1]))
        if n_layers_sub2 > 1:
            self._factor_sub2 *= int(np.prod(subsamples[:n_layers_sub2 - 1]))
        self._factor *= int(np.prod(subsamples))
        for n, p in self.named_parameters():       # reference :256-262: uniform(-param_init, param_init), biases 0
            if 'conv' in n:                       # CNN front-end (and a conv1d subsampler) keep their own init
                continue
            if p.dim() == 1:
                nn.init.constant_(p, 0.)
            else:
                nn.init.uniform_(p, a=-param_init, b=param_init)
        self.reset_cache()

    def reset_cache(self):
        self.hx_fwd = [None] * self.n_layers
        self.hx_bwd = [None] * self.n_layers

    def _run_lstm(self, rnn, tag, xs, lens_dev, names, state=None, want_state=False):
        """Input projection of all frames (one GEMM, both biases folded in) + the persistent recurrence kernel for the
        directions in `names` of `rnn`.  state / want_state: (h, c) `[n_dirs, B, H]` carried across chunks."""
        prec = get_precision(self)
        w_ih = [getattr(rnn, 'weight_ih' + n) for n in names]
        w_ihp = prepared(self, 'w_ih' + tag, prec, tuple(w_ih), build=lambda *ws: torch.cat(ws, dim=0))
        bias = cached(self, 'b' + tag, tuple(getattr(rnn, 'bias_ih' + n) for n in names) +
                      tuple(getattr(rnn, 'bias_hh' + n) for n in names),
                      lambda *bs: (torch.cat(bs[:len(names)]) + torch.cat(bs[len(names):])).float().contiguous())
        w_hh = cached(self, 'w_hh' + tag, tuple(getattr(rnn, 'weight_hh' + n) for n in names),
                      lambda *ws: torch.stack(ws, dim=0).float().contiguous())
        gates_x = ops.linear(xs, w_ihp, bias, prec=prec, out_dtype=torch.float32)
        if state is not None or want_state:
            return ops.lstm_seq(gates_x, w_hh, lens_dev, len(names), state=state, want_state=True, prec=prec)
        return ops.lstm_seq(gates_x, w_hh, lens_dev, len(names), prec=prec)

    def _lstm_layer(self, lth, xs, lens_dev, train=False, state=None, want_state=False):
        """One (bi)directional LSTM layer over `[B, T, I]` with packed-sequence semantics."""
        prec = get_precision(self)
        if train:
            ys = ag.lstm_layer(self, lth, xs, lens_dev, prec)
        else:
            names = ['_l0'] + (['_l0_reverse'] if self.bidirectional else [])
            ys = self._run_lstm(self.rnn[lth], '%d' % lth, xs, lens_dev, names, state, want_state)
            if state is not None or want_state:
                ys, state = ys
        if self.bidir_sum and self.bidirectional:
            half = ys.size(-1) // 2
            ys = ys[:, :, :half] + ys[:, :, half:]
        return (ys, state) if want_state else ys

    def _full_lens(self, xs):
        return lens_to_device(torch.IntTensor([xs.size(1)] * xs.size(0)), xs.device)

    def _uni(self, rnn, tag, xs, state, train):
        """One unidirectional LSTM over all frames of `xs` from `state` -> (ys, new state)."""
        full = self._full_lens(xs)
        if train:
            return ag.lstm_chunk(self, tag, rnn, xs, full, state, get_precision(self))
        return self._run_lstm(rnn, tag, xs, full, ['_l0'], state, want_state=True)

    def _lc_layer(self, lth, xs, n_carry, train=False):
        """One latency-controlled layer over a chunk `[B, <= N_c + N_r, I]` (reference :460-481): the backward LSTM runs
        over the whole chunk from a zero state; the forward LSTM continues from the carried state, which is saved after
        the first `n_carry` (= N_c at this depth) frames so that the look-ahead frames do not leak into the next chunk.
        In training the carried state stays attached to the graph: the loss back-propagates through it into earlier chunks."""
        ys_bwd = torch.flip(self._uni(self.rnn_bwd[lth], 'bwd%d' % lth, torch.flip(xs, dims=[1]), None, train)[0], dims=[1])
        if xs.size(1) <= n_carry:                                   # last chunk of the utterance
            ys_fwd, self.hx_fwd[lth] = self._uni(self.rnn[lth], '%d' % lth, xs, self.hx_fwd[lth], train)
        else:
            head, tail = xs[:, :n_carry].contiguous(), xs[:, n_carry:].contiguous()
            y1, self.hx_fwd[lth] = self._uni(self.rnn[lth], '%d' % lth, head, self.hx_fwd[lth], train)
            y2, _ = self._uni(self.rnn[lth], '%d' % lth, tail, self.hx_fwd[lth], train)
            ys_fwd = torch.cat([y1, y2], dim=1)
        ys = ys_fwd + ys_bwd if self.bidir_sum else torch.cat([ys_fwd, ys_bwd], dim=-1)
        return ag.dropout(ys, self.dropout.p) if train else ys

    def _lc_tail(self, lth, xs, xlens, train=False):
        """Projection (+ReLU) and subsampling after an LC layer; returns (xs, xlens)."""
        prec = get_precision(self)
        if self.proj is not None and lth != self.n_layers - 1:
            lin = self.proj[lth]
            if train:
                xs = ag.linear_relu(self, 'proj%d' % lth, lin.weight, lin.bias, xs, prec)
            else:
                xs = ops.linear(xs, prepared(self, 'proj%d' % lth, prec, (lin.weight,)), lin.bias, prec=prec, act='relu')
        if self.subsample is not None:
            xs, xlens = ag.subsample_train(self.subsample[lth], xs, xlens) if train else self.subsample[lth](xs, xlens)
        return xs, xlens

    def _forward_full_context(self, xs, xlens, train=False):
        """LC encoder without a current-chunk size (N_c <= 0): whole utterance at once (reference :385-425)."""
        xs_sub1 = xlens_sub1 = None
        for lth in range(self.n_layers):
            ys_bwd = torch.flip(self._uni(self.rnn_bwd[lth], 'bwd%d' % lth, torch.flip(xs, dims=[1]), None, train)[0], dims=[1])
            ys_fwd, self.hx_fwd[lth] = self._uni(self.rnn[lth], '%d' % lth, xs, self.hx_fwd[lth], train)
            xs = ys_fwd + ys_bwd if self.bidir_sum else torch.cat([ys_fwd, ys_bwd], dim=-1)
            if train:
                xs = ag.dropout(xs, self.dropout.p)
            if lth == self.n_layers_sub1 - 1:
                xs_sub1, xlens_sub1 = self._sub_out(xs, 'sub1', train), xlens.clone()
            xs, xlens = self._lc_tail(lth, xs, xlens, train)
        return xs, xlens, xs_sub1, xlens_sub1

    def _forward_latency_controlled(self, xs, xlens, N_c, N_r, streaming, train=False):
        """Chunk loop of the LC-BLSTM (reference :427-510): layer loop inside the chunk loop; streaming = one chunk."""
        bs, xmax, _ = xs.size()
        n_chunks = math.ceil(xmax / N_c)
        if streaming:
            xlens = torch.IntTensor(bs).fill_(min(xmax, N_c))
        xlens_sub1 = xlens.clone() if self.n_layers_sub1 > 0 else None
        chunks, chunks_sub1 = [], []
        for chunk_idx, t in enumerate(range(0, N_c * n_chunks, N_c)):
            xs_chunk = xs[:, t:t + (N_c + N_r)].contiguous()
            n_c = N_c
            for lth in range(self.n_layers):
                xs_chunk = self._lc_layer(lth, xs_chunk, n_c, train)
                if lth == self.n_layers_sub1 - 1:
                    chunks_sub1.append(xs_chunk[:, :n_c].clone())
                    if chunk_idx == 0:
                        xlens_sub1 = xlens.clone()
                xs_chunk, xlens_tmp = self._lc_tail(lth, xs_chunk, xlens, train)
                if self.subsample is not None:
                    if chunk_idx == 0:
                        xlens = xlens_tmp
                    n_c = n_c // self.subsample[lth].factor
            chunks.append(xs_chunk[:, :n_c])
            if streaming:
                break
        xs = torch.cat(chunks, dim=1)
        xs_sub1 = None
        if self.n_layers_sub1 > 0:
            xs_sub1 = self._sub_out(torch.cat(chunks_sub1, dim=1), 'sub1', train)
        return xs, xlens, xs_sub1, xlens_sub1

    def _sub_out(self, xs, module, train=False):
        prec = get_precision(self)
        if train:
            xs_sub = xs.clone()
            if self.task_specific_layer:
                lin = getattr(self, 'layer_' + module)
                xs_sub = ag.dropout(ag.linear_relu(self, 'layer_' + module, lin.weight, lin.bias, xs, prec), self.dropout.p)
            bridge = getattr(self, 'bridge_' + module)
            return xs_sub if bridge is None else ag.linear(self, 'bridge_' + module, bridge, xs_sub, prec)
        if self.task_specific_layer:
            lin = getattr(self, 'layer_' + module)
            xs_sub = ops.linear(xs, prepared(self, 'layer_' + module, prec, (lin.weight,)), lin.bias, prec=prec, act='relu')
        else:
            xs_sub = xs.clone()
        bridge = getattr(self, 'bridge_' + module)
        if bridge is not None:
            xs_sub = ops.linear(xs_sub, prepared(self, 'bridge_' + module, prec, (bridge.weight,)), bridge.bias, prec=prec)
        return xs_sub

    def _subsample_train(self, lth, xs, xlens):
        return ag.subsample_train(self.subsample[lth], xs, xlens)

    def forward(self, xs, xlens, task, streaming=False, lookback=False, lookahead=False):
        if self.training and not ag.training_enabled(self) and (
                self.dropout_in.p > 0 or (self.enc_type != 'conv' and self.dropout.p > 0)):
            raise NotImplementedError("dropout > 0 in train() mode needs the autograd training path (grad enabled)")
        eouts = {'ys': {'xs': None, 'xlens': None}, 'ys_sub1': {'xs': None, 'xlens': None},
                 'ys_sub2': {'xs': None, 'xlens': None}}
        xlens = torch.IntTensor([int(v) for v in xlens])
        prec = get_precision(self)
        # train() + grad mode: autograd nodes with hand-written CUDA backward; otherwise inference kernels under no_grad
        train = ag.training_enabled(self)
        if train and self.rsp_prob > 0:
            raise NotImplementedError("random state passing (rsp_prob > 0) needs a carried initial state in the training "
                                      "LSTM node; not on the B200 path (it is inactive in eval mode)")
        if train and streaming:
            raise NotImplementedError("streaming=True is an inference path (call .eval() / torch.no_grad())")
        with (torch.enable_grad() if train else torch.no_grad()):
            bs = xs.size(0)
            N_c, N_r = self.N_c, self.N_r
            if train:                                                     # dropout_in (:301)
                xs = ag.dropout(xs.float().contiguous(), self.dropout_in.p)
            if self.lc_bidir and not self.cnn_lookahead:                  # CNN applied chunk by chunk (:308-312)
                xs = chunkwise(xs, 0, N_c, 0).contiguous().view(bs, -1, xs.size(2))[:, :int(xlens.max())]
            if self.conv is not None:
                if train:
                    if lookback or lookahead:
                        raise NotImplementedError("CNN lookback/lookahead trimming is an inference (streaming) feature")
                    if self.conv.is_1dconv:
                        xs, xlens = self.conv(xs, xlens)
                    else:
                        xs, xlens = ag.frontend_forward(self.conv, xs, 1.0, prec), self.conv.output_lens(xlens)
                else:
                    xs, xlens = self.conv(xs, xlens, lookback=lookback, lookahead=lookahead)
                if self.enc_type == 'conv':
                    eouts['ys']['xs'], eouts['ys']['xlens'] = xs, xlens
                    return eouts
                if self.lc_bidir:
                    N_c, N_r = N_c // self.conv_factor, N_r // self.conv_factor
            if not streaming:
                self.reset_cache()
            xs = xs.float()
            if self.lc_bidir:
                if self.N_c <= 0:
                    xs, xlens, xs_sub1, xlens_sub1 = self._forward_full_context(xs, xlens, train)
                else:
                    xs, xlens, xs_sub1, xlens_sub1 = self._forward_latency_controlled(xs, xlens, N_c, N_r, streaming, train)
                if task == 'ys_sub1':
                    eouts[task]['xs'], eouts[task]['xlens'] = xs_sub1, xlens_sub1
                    return eouts
            for lth in range(self.n_layers if not self.lc_bidir else 0):
                if streaming:       # un-packed `rnn(xs, hx=prev_state)` (:541): every frame of the chunk, state carried
                    xs, self.hx_fwd[lth] = self._lstm_layer(lth, xs, self._full_lens(xs), state=self.hx_fwd[lth],
                                                            want_state=True)
                else:
                    xs = self._lstm_layer(lth, xs, lens_to_device(xlens, xs.device), train)
                    if train:                                             # dropout after every layer (:347)
                        xs = ag.dropout(xs, self.dropout.p)
                if lth == self.n_layers_sub1 - 1:
                    xs_sub1, xlens_sub1 = self._sub_out(xs, 'sub1', train), xlens.clone()
                    if task == 'ys_sub1':
                        eouts[task]['xs'], eouts[task]['xlens'] = xs_sub1, xlens_sub1
                        return eouts
                if lth == self.n_layers_sub2 - 1:
                    xs_sub2, xlens_sub2 = self._sub_out(xs, 'sub2', train), xlens.clone()
                    if task == 'ys_sub2':
                        eouts[task]['xs'], eouts[task]['xlens'] = xs_sub2, xlens_sub2
                        return eouts
                if self.proj is not None and lth != self.n_layers - 1:
                    lin = self.proj[lth]
                    if train:
                        xs = ag.linear_relu(self, 'proj%d' % lth, lin.weight, lin.bias, xs, prec)
                    else:
                        xs = ops.linear(xs, prepared(self, 'proj%d' % lth, prec, (lin.weight,)), lin.bias, prec=prec, act='relu')
                if self.subsample is not None:
                    xs, xlens = self._subsample_train(lth, xs, xlens) if train else self.subsample[lth](xs, xlens)
            if self.bridge is not None:
                if train:
                    xs = ag.linear(self, 'bridge', self.bridge, xs, prec)
                else:
                    xs = ops.linear(xs, prepared(self, 'bridge', prec, (self.bridge.weight,)), self.bridge.bias, prec=prec)
            xs = xs[:, :int(xlens.max())]
        if task in ['all', 'ys']:
            eouts['ys']['xs'], eouts['ys']['xlens'] = xs, xlens
        if self.n_layers_sub1 >= 1 and task == 'all':
            eouts['ys_sub1']['xs'], eouts['ys_sub1']['xlens'] = xs_sub1, xlens_sub1
        if self.n_layers_sub2 >= 1 and task == 'all':
            eouts['ys_sub2']['xs'], eouts['ys_sub2']['xlens'] = xs_sub2, xlens_sub2
        return eouts


# command-line contract of the reference (add_args / define_name static methods): see encoders/cli.py
from . import cli as _cli  # noqa: E402

RNNEncoder.add_args = staticmethod(_cli.rnn_add_args)
RNNEncoder.define_name = staticmethod(_cli.rnn_define_name)
