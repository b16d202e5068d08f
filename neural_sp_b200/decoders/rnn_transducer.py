"""RNN-Transducer decoder, training path (reference decoders/rnn_transducer.py:32-311), B200-native.

Same constructor and parameter names (``rnn.N``, ``embed``, ``w_enc``, ``w_dec``, ``output``, ``ctc.*``).
``forward`` / ``forward_transducer`` compute the loss: prediction network (embedding gather + stacked LSTMs on the
library's persistent LSTM kernel, the nn.LSTM modules only hold the parameters), joint network (two GEMMs, fused add+tanh,
vocabulary GEMM, row log-softmax) and the RNN-T lattice kernel, which yields the loss and d loss / d log_probs in one
pass.  In train() + grad mode the whole chain is differentiable through hand-written backward kernels
(neural_sp_b200/autograd.py: _RnntJointLossFn, _LstmLayerFn, _LinearFn, _LinearReluFn, _DropoutFn), so the loss
back-propagates into the encoder output, the prediction network and the embedding; only the embedding row gather /
scatter-add is torch's (data movement).  Beam search (:419-819) is out of scope."""
import copy

import torch
import torch.nn as nn

from .. import autograd as ag
from .. import ops
from ..modules._prep import prepared, cached, get_precision, act_dtype
from .ctc import CTC, _lens_dev


class RNNTransducer(nn.Module):
    def __init__(self, special_symbols, enc_n_units, n_units, n_projs, n_layers, bottleneck_dim, emb_dim, vocab,
                 dropout, dropout_emb, ctc_weight, ctc_lsm_prob, ctc_fc_list, external_lm, global_weight,
                 mtl_per_batch, param_init):
        super().__init__()
        self.eos = special_symbols['eos']
        self.unk = special_symbols['unk']
        self.pad = special_symbols['pad']
        self.blank = special_symbols['blank']
        self.vocab = vocab
        self.enc_n_units = enc_n_units
        self.dec_n_units = n_units
        self.n_projs = n_projs
        self.n_layers = n_layers
        self.rnnt_weight = global_weight - ctc_weight
        self.ctc_weight = ctc_weight
        self.mtl_per_batch = mtl_per_batch
        self.prev_spk = ''
        self.lmstate_final = None
        self.embed_cache = None
        self.bidirectional = False             # read by the shared LSTM layer node (autograd.lstm_layer)
        if external_lm is not None:
            raise NotImplementedError("LM initialisation of the prediction network is out of scope")
        if ctc_weight > 0:
            self.ctc = CTC(eos=self.eos, blank=self.blank, enc_n_units=enc_n_units, vocab=vocab, dropout=dropout,
                           lsm_prob=ctc_lsm_prob, fc_list=ctc_fc_list, param_init=0.1)
        if self.rnnt_weight > 0:
            self.rnn = nn.ModuleList()
            dec_odim = emb_dim
            # one Linear deep-copied n_layers times, like the reference's repeat(): same RNG consumption, so that
            # seeded construction yields bit-identical initial weights
            proto = nn.Linear(n_units, n_projs) if n_projs > 0 else None
            self.proj = nn.ModuleList([copy.deepcopy(proto) for _ in range(n_layers)]) if n_projs > 0 else None
            self.dropout = nn.Dropout(p=dropout)
            for _ in range(n_layers):
                self.rnn += [nn.LSTM(dec_odim, n_units, 1, batch_first=True)]
                dec_odim = n_projs if n_projs > 0 else n_units
            self.embed = nn.Embedding(vocab, emb_dim, padding_idx=self.pad)
            self.dropout_emb = nn.Dropout(p=dropout_emb)
            self.w_enc = nn.Linear(enc_n_units, bottleneck_dim)
            self.w_dec = nn.Linear(dec_odim, bottleneck_dim, bias=False)
            self.output = nn.Linear(bottleneck_dim, vocab)
        for n, p in self.named_parameters():       # reference :161-172: uniform(-param_init, param_init), biases 0,
            if p.dim() == 1:                       # INCLUDING the auxiliary CTC head (no exclusion in the reference)
                nn.init.constant_(p, 0.)
            else:
                nn.init.uniform_(p, a=-param_init, b=param_init)

    def set_precision(self, precision):
        for m in self.modules():
            m.precision = precision
        return self

    def _lstm(self, lth, xs, lens_dev):
        """One prediction-network LSTM layer (inference kernels): input GEMM + persistent recurrence."""
        rnn, prec = self.rnn[lth], get_precision(self)
        w_ihp = prepared(self, 'w_ih%d' % lth, prec, (rnn.weight_ih_l0,))
        bias = cached(self, 'b%d' % lth, (rnn.bias_ih_l0, rnn.bias_hh_l0), lambda a, b: (a + b).float().contiguous())
        w_hh = cached(self, 'w_hh%d' % lth, (rnn.weight_hh_l0,), lambda w: w.unsqueeze(0).float().contiguous())
        return ops.lstm_seq(ops.linear(xs, w_ihp, bias, prec=prec, out_dtype=torch.float32), w_hh, lens_dev, 1, prec=prec)

    def recurrency(self, ys_emb, train=False):
        """Prediction network (reference :278-311), zero initial state: stacked unidirectional LSTMs over all U+1 positions
        (the reference does not pack here), dropout, optional projection + ReLU."""
        prec = get_precision(self)
        B, U1, _ = ys_emb.shape
        lens = _lens_dev(torch.IntTensor([U1] * B), ys_emb.device)
        out = ys_emb
        for lth in range(self.n_layers):
            if train:
                out = ag.dropout(ag.lstm_layer(self, lth, out, lens, prec), self.dropout.p)
            else:
                out = self._lstm(lth, out.float(), lens)
            if self.proj is not None:
                lin = self.proj[lth]
                if train:
                    out = ag.linear_relu(self, 'proj%d' % lth, lin.weight, lin.bias, out, prec)
                else:
                    out = ops.linear(out, prepared(self, 'proj%d' % lth, prec, (lin.weight,)), lin.bias, prec=prec, act='relu')
        return out

    def joint(self, eouts, dout):
        """log_probs `[B, T, U+1, vocab]` = log_softmax(output(tanh(w_enc(e)[:, :, None] + w_dec(d)[:, None]))) (:262-276, :242)."""
        prec = get_precision(self)
        B, T, _ = eouts.shape
        U1 = dout.shape[1]
        e = ops.linear(eouts, prepared(self, "w_enc", prec, (self.w_enc.weight,)), self.w_enc.bias, prec=prec)
        d = ops.linear(dout, prepared(self, "w_dec", prec, (self.w_dec.weight,)), None, prec=prec)
        h = ops.rnnt_joint_tanh(e, d, out_dtype=act_dtype(prec))
        logits = ops.linear(h.view(B * T * U1, -1), prepared(self, "output", prec, (self.output.weight,)),
                            self.output.bias, prec=prec, out_dtype=torch.float32)
        return ops.softmax_rows(logits.view(B, T, U1, self.vocab), log=True, inplace=True)

    def _pack(self, ys, device):
        """Label lists -> (prediction-network input `[B, U+1]` long: <eos> + labels, padded; targets int32 `[B, U]`; lengths
        int32 `[B]`), all on `device` (reference :225-231).  Host buffers are pinned and the last batch is cached, as
        ops.pack_labels does for CTC: a repeated batch (CUDA-graph replay) issues no new copy."""
        key = (tuple(tuple(int(v) for v in y) for y in ys), str(device))
        hit = self.__dict__.get("_pack_cache")
        if hit is not None and hit[0] == key:
            return hit[1]
        B = len(ys)
        ylens = [len(y) for y in ys]
        U = max(ylens) if ylens else 0
        ys_in = torch.full((B, U + 1), self.pad, dtype=torch.long)
        ys_out = torch.zeros((B, max(U, 1)), dtype=torch.int32)
        for b, y in enumerate(ys):
            ys_in[b, 0] = self.eos
            if len(y):
                ys_in[b, 1:len(y) + 1] = torch.as_tensor(list(y), dtype=torch.long)
                ys_out[b, :len(y)] = torch.as_tensor(list(y), dtype=torch.int32)
        labels = ys_out[:, :U].contiguous()
        if device.type == "cuda":
            ys_in, labels = ys_in.pin_memory(), labels.pin_memory()
        out = (ys_in.to(device, non_blocking=True),
               labels.to(device, non_blocking=True) if U > 0 else torch.zeros(B, 0, dtype=torch.int32, device=device),
               _lens_dev(torch.tensor(ylens, dtype=torch.int32), device))
        self.__dict__["_pack_cache"] = (key, out)
        return out

    def forward_transducer(self, eouts, elens, ys):
        """RNN-T loss (reference :217-260): mean over the batch of -log p(y|x); returns a `[1]` tensor (differentiable in
        train() + grad mode)."""
        device = eouts.device
        B = len(ys)
        ys_in, labels, ylens_d = self._pack(ys, device)
        flens = _lens_dev(elens, device)
        prec = get_precision(self)
        if self.training and torch.is_grad_enabled():
            emb = torch.nn.functional.embedding(ys_in, self.embed.weight, padding_idx=self.pad)
            dout = self.recurrency(ag.dropout(emb.float(), self.dropout_emb.p), train=True)
            e = ag.linear(self, 'w_enc', self.w_enc, eouts.float(), prec)                  # `[B, T, J]`
            d = ag.linear(self, 'w_dec', self.w_dec, dout, prec)                           # `[B, U+1, J]`
            loss, _ = ag.rnnt_joint_loss(self, e, d, labels, flens, ylens_d, self.blank, prec)
            self._grad_log_probs = None
            return loss.reshape(1)
        with torch.no_grad():
            dout = self.recurrency(self.embed(ys_in))
            log_probs = self.joint(eouts.float(), dout.float())
            loss, nll, grad = ops.rnnt_loss_fwd_bwd(log_probs, labels, flens, ylens_d, self.blank, need_grad=True)
        self._grad_log_probs = grad            # d loss / d log_probs, kept for callers that chain the backward by hand
        return loss.reshape(1)

    def forward(self, eouts, elens, ys, task='all', teacher_logits=None, recog_params={}, idx2token=None, trigger_points=None):
        """Reference :174-215: total loss = ctc_weight * CTC + rnnt_weight * RNN-T; observation dict of floats."""
        observation = {'loss': None, 'loss_transducer': None, 'loss_ctc': None, 'loss_mbr': None}
        loss = eouts.new_zeros((1,))
        if self.ctc_weight > 0 and (task == 'all' or 'ctc' in task):
            loss_ctc, _ = self.ctc(eouts, elens, ys)
            observation['loss_ctc'] = float(loss_ctc.detach())
            loss = loss + (loss_ctc * (1 if self.mtl_per_batch else self.ctc_weight)).reshape(1)
        if self.rnnt_weight > 0 and (task == 'all' or 'ctc' not in task):
            loss_t = self.forward_transducer(eouts, elens, ys)
            observation['loss_transducer'] = float(loss_t.detach())
            loss = loss + loss_t * (1 if self.mtl_per_batch else self.rnnt_weight)
        observation['loss'] = float(loss.detach())
        return loss, observation


# command-line contract (reference decoders/rnn_transducer.py:130-158)
def _rnnt_add_args(parser, args):
    group = parser.add_argument_group("RNN-T decoder")
    if not hasattr(args, 'dec_n_units'):           # shared with the LAS decoder's options
        group.add_argument('--dec_n_units', type=int, default=512)
        group.add_argument('--dec_n_projs', type=int, default=0)
        group.add_argument('--dec_bottleneck_dim', type=int, default=1024)
        group.add_argument('--emb_dim', type=int, default=512)
    return parser


def _rnnt_define_name(dir_name, args):
    name = dir_name + '_' + args.dec_type + '%dH' % args.dec_n_units
    if args.dec_n_projs > 0:
        name += '%dP' % args.dec_n_projs
    return name + '%dL' % args.dec_n_layers


RNNTransducer.add_args = staticmethod(_rnnt_add_args)
RNNTransducer.define_name = staticmethod(_rnnt_define_name)
