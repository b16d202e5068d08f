"""CTC head + loss + forced aligner, B200-native.

Drop-in for ``neural_sp.models.seq2seq.decoders.ctc.CTC`` (reference ctc.py:35-150) and
``CTCForcedAligner`` (ctc.py:628-753): same constructor, ``forward(eouts, elens, ys, forced_align)``
-> ``(loss, trigger_points)``, ``loss_fn(logits[T,B,V], ys_ctc, elens, ylens)``, same ``state_dict``
keys (``output.weight/bias`` or ``output.fc{i}.weight/bias``).  The arithmetic runs in
libnsp_b200.so: one fused CUDA pass computes the loss AND d(loss)/d(logits); ``backward`` only scales.

Decode helpers: greedy / trigger points on the device (decode.cu); prefix beam search, offline and block-synchronous
(ctc.py:245-531, the decoder half of streaming), in ctc_beam.py on top of the log-softmax kernel.  The attention decoders'
prefix scorer (ctc.py:756-871) is out of scope.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from . import ctc_beam
from ..modules.dropout import Dropout
from ..modules.linear import Linear


def _lens_dev(elens, device):
    from ..encoders.transformer import lens_to_device
    return lens_to_device(elens, device)


class _CTCLossFn(torch.autograd.Function):
    """loss = (1-lsm) * sum_b nll_b / B + lsm * KL ; gradient produced by the same kernels."""

    @staticmethod
    def forward(ctx, logits_btv, labels, elens, ylens, blank, lsm_prob):
        loss, nll, grad = ops.ctc_loss_fwd_bwd(logits_btv, labels, elens, ylens, blank, lsm_prob)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(nll)
        return loss, nll

    @staticmethod
    def backward(ctx, g_loss, g_nll):
        (grad,) = ctx.saved_tensors
        return grad * g_loss, None, None, None, None, None


def ctc_loss(logits_btv, labels, elens, ylens, blank=0, lsm_prob=0.0):
    """Functional form on device-resident int32 labels/lengths. Returns (loss, nll)."""
    return _CTCLossFn.apply(logits_btv, labels, elens, ylens, blank, lsm_prob)


class CTCForcedAligner(object):
    """Reference ctc.py:628-753; integer trigger points are bit-exact targets."""

    def __init__(self, blank=0):
        self.blank = blank

    def __call__(self, logits, elens, ys, ylens=None):
        """logits `[B, T, vocab]` (CUDA), elens IntTensor `[B]`, ys list of lists -> IntTensor `[B, Lmax+1]`."""
        with torch.no_grad():
            labels, ylens_d, _ = ops.pack_labels(ys, logits.device)
            elens_d = elens.to(device=logits.device, dtype=torch.int32, non_blocking=True)
            return ops.ctc_forced_align(logits.detach().float(), labels, elens_d, ylens_d, self.blank)


class _BeamParams:
    """What `_beam_search` reads from the reference's BeamSearch helper (beam_search.py:19-31)."""

    def __init__(self, beam_width, lm_weight):
        self.beam_width, self.lm_weight = beam_width, lm_weight


class CTC(nn.Module):
    """Connectionist temporal classification (reference ctc.py:35-137)."""

    def __init__(self, eos, blank, enc_n_units, vocab, dropout=0., lsm_prob=0., fc_list=None,
                 param_init=0.1, backward=False):
        super().__init__()
        self.eos = eos
        self.blank = blank
        self.vocab = vocab
        self.lsm_prob = lsm_prob
        self.bwd = backward
        self.space = -1
        self.prev_spk = ''
        self.lmstate_final = None
        self.prob_dict = {}
        self.data_dict = {}

        # Fully-connected layers before the softmax (no nonlinearity in between: ctc.py:82-89)
        if fc_list is not None and len(fc_list) > 0:
            _fc_list = [int(fc) for fc in fc_list.split('_')]
            fc_layers = OrderedDict()
            for i in range(len(_fc_list)):
                input_dim = enc_n_units if i == 0 else _fc_list[i - 1]
                fc_layers['fc' + str(i)] = Linear(input_dim, _fc_list[i])
                fc_layers['dropout' + str(i)] = Dropout(p=dropout)
            fc_layers['fc' + str(len(_fc_list))] = Linear(_fc_list[-1], vocab)
            self.output = nn.Sequential(fc_layers)
        else:
            self.output = Linear(enc_n_units, vocab)
        self.forced_aligner = CTCForcedAligner(blank)

    def set_precision(self, precision):
        """'bf16' | 'tf32' | 'fp32' arithmetic of the output head's GEMMs (the loss itself is always fp32)."""
        for m in self.modules():
            m.precision = precision
        return self

    def forward(self, eouts, elens, ys, forced_align=False):
        """Compute CTC loss.

        Args:
            eouts (FloatTensor): `[B, T, enc_n_units]` on the GPU
            elens (IntTensor): `[B]`
            ys (List): length `[B]`, each a list of label ids
        Returns:
            loss (FloatTensor): 0-dim
            trigger_points (IntTensor): `[B, L+1]` or None
        """
        ys_dir = [list(y[::-1]) if self.bwd else list(y) for y in ys]
        labels, ylens_d, _ = ops.pack_labels(ys_dir, eouts.device)
        elens_d = _lens_dev(elens, eouts.device)      # cached pinned copy: no sync, CUDA-graph safe

        logits = self.output(eouts)   # `[B, T, vocab]`
        loss, _ = ctc_loss(logits.float(), labels, elens_d, ylens_d, self.blank, self.lsm_prob)

        trigger_points = None
        if forced_align:
            with torch.no_grad():
                labels_f, ylens_f, _ = ops.pack_labels([list(y) for y in ys], eouts.device)
                trigger_points = ops.ctc_forced_align(logits.detach().float(), labels_f, elens_d, ylens_f, self.blank)

        if not self.training:
            self.data_dict['elens'] = elens.cpu().numpy() if torch.is_tensor(elens) else np.asarray(elens)
            self.prob_dict['probs'] = torch.softmax(logits.detach(), dim=-1).cpu().numpy()
        return loss, trigger_points

    def probs(self, eouts, temperature=1.):
        """CTC probabilities `[B, T, vocab]` (reference ctc.py:197-206)."""
        with torch.no_grad():
            return ops.softmax_rows(self.output(eouts).float(), log=False, temperature=temperature)

    def scores(self, eouts, temperature=1.):
        """Log-scale CTC probabilities `[B, T, vocab]` (reference ctc.py:208-217)."""
        with torch.no_grad():
            return ops.softmax_rows(self.output(eouts).float(), log=True, temperature=temperature)

    # ---- prefix beam search (host bookkeeping in ctc_beam.py; the frame scores come from the kernels behind `scores`) ----
    def initialize_beam(self, hyp, lmstate):
        return ctc_beam.initialize_beam(hyp, lmstate)

    def _frame_scores(self, eouts, softmax_smoothing):
        """log_softmax(output(eouts) * softmax_smoothing) `[B, T, vocab]` (reference ctc.py:299 / :509)."""
        with torch.no_grad():
            logits = self.output(eouts).float()
            if softmax_smoothing != 1.0:
                logits = logits * softmax_smoothing
            return ops.softmax_rows(logits, log=True)

    def _beam_search(self, hyps, helper, scores_ctc, lm, lp_weight):
        """`scores_ctc` `[T, vocab]` log-probabilities; `helper` supplies `beam_width` and `lm_weight` (reference :365-483)."""
        if not hasattr(self, 'state_cache'):
            self.state_cache = OrderedDict()
        block = ctc_beam.frame_scores_to_host(scores_ctc)
        return ctc_beam.prefix_beam_search(hyps, block, helper.beam_width, self.vocab, self.blank, lm, helper.lm_weight,
                                           lp_weight, self.state_cache)

    def beam_search(self, eouts, elens, params, idx2token, lm=None, lm_second=None, lm_second_bwd=None,
                    nbest=1, refs_id=None, utt_ids=None, speakers=None):
        """Offline prefix beam search (reference ctc.py:256-363) -> N-best token-id arrays per utterance (without <eos>).
        Like the reference, every frame of `eouts` is consumed (`elens` is not used to cut the padding).  First-pass LM
        shallow fusion is supported through the LM's `predict`; second-pass rescoring LMs are outside this package."""
        if lm_second is not None or lm_second_bwd is not None:
            raise NotImplementedError("second-pass LM rescoring (reference beam_search.py:116-141) is not part of the B200 path")
        beam_width = params.get('recog_beam_width')
        lp_weight = params.get('recog_length_penalty')
        lm_weight = params.get('recog_lm_weight')
        lm_state_CO = params.get('recog_lm_state_carry_over')
        softmax_smoothing = params.get('recog_softmax_smoothing')
        if lm is not None:                                  # reference beam_search.py:143-149
            assert lm_weight > 0
            lm.eval()
            if params.get('recog_cache_embedding'):
                lm.cache_embedding(lm.device)
        helper = _BeamParams(beam_width, lm_weight)
        log_probs = self._frame_scores(eouts, softmax_smoothing if softmax_smoothing is not None else 1.0)
        nbest_hyps_idx, end_hyps = [], []
        for b in range(eouts.size(0)):
            lmstate = {'hxs': eouts.new_zeros(lm.n_layers, 1, lm.n_units),
                       'cxs': eouts.new_zeros(lm.n_layers, 1, lm.n_units)} if lm is not None else None
            if speakers is not None:
                if speakers[b] == self.prev_spk and lm_state_CO:
                    lmstate = self.lmstate_final
                self.prev_spk = speakers[b]
            self.state_cache = OrderedDict()
            hyps, new_hyps = self._beam_search(self.initialize_beam([self.eos], lmstate), helper, log_probs[b], lm, lp_weight)
            end_hyps = hyps[:]
            if len(end_hyps) < nbest and nbest > 1:
                end_hyps.extend(new_hyps[:nbest - len(end_hyps)])
            end_hyps = sorted(end_hyps, key=lambda x: x['score'] / max(len(x['hyp'][1:]), 1), reverse=True)   # length-normalised
            nbest_hyps_idx += [[np.array(end_hyps[n]['hyp'][1:]) for n in range(nbest)]]
        if eouts.size(0) == 1:
            self.lmstate_final = end_hyps[0]['lmstate']
        return nbest_hyps_idx

    def beam_search_block_sync(self, eouts, params, helper, idx2token, hyps, lm):
        """One block of streaming decoding (reference ctc.py:485-531; caller speech2text.py:627): `eouts` `[1, T_block, D]`,
        `hyps` = what the previous block returned (None at the start of an utterance) -> (end_hyps = [], hyps)."""
        assert eouts.size(0) == 1
        beam_width = params.get('recog_beam_width')
        lp_weight = params.get('recog_length_penalty')
        lm_state_CO = params.get('recog_lm_state_carry_over')
        softmax_smoothing = params.get('recog_softmax_smoothing')
        end_hyps = []
        if hyps is None:
            if lm_state_CO:
                lmstate = self.lmstate_final
            else:
                lmstate = {'hxs': eouts.new_zeros(lm.n_layers, 1, lm.n_units),
                           'cxs': eouts.new_zeros(lm.n_layers, 1, lm.n_units)} if lm is not None else None
            self.n_frames = 0
            hyps = self.initialize_beam([self.eos], lmstate)
            self.state_cache = OrderedDict()
        log_probs = self._frame_scores(eouts, softmax_smoothing if softmax_smoothing is not None else 1.0)
        hyps, _ = self._beam_search(hyps, helper, log_probs[0], lm, lp_weight)
        merged_hyps = sorted(end_hyps + hyps, key=lambda x: x['score'], reverse=True)[:beam_width]
        if len(merged_hyps) > 0:
            self.lmstate_final = merged_hyps[0]['lmstate']
        self.n_frames += eouts.size(1)
        return end_hyps, hyps

    def _greedy_device(self, eouts, elens):
        with torch.no_grad():
            logits = self.output(eouts).float()
            elens_d = _lens_dev(torch.as_tensor(np.asarray(elens), dtype=torch.int32), eouts.device)
            return ops.ctc_greedy(logits, elens_d, self.blank)

    def greedy(self, eouts, elens):
        """Greedy decoding (reference ctc.py:219-243): argmax path, repeats collapsed, blanks removed.
        Returns ``[[hyp_b]]`` per utterance like the reference."""
        _, hyp, hyp_lens, _ = self._greedy_device(eouts, elens)
        hyp, hyp_lens = hyp.cpu().numpy(), hyp_lens.cpu().numpy()
        return [[hyp[b, :hyp_lens[b]].tolist()] for b in range(eouts.size(0))]

    def trigger_points(self, eouts, elens):
        """Trigger points of the greedy path (reference ctc.py:152-195): IntTensor `[B, Lmax+1]`."""
        _, _, hyp_lens, trig = self._greedy_device(eouts, elens)
        ymax = int(hyp_lens.max().item())
        out = torch.zeros(eouts.size(0), ymax + 1, dtype=torch.int32, device=eouts.device)
        if ymax > 0:
            mask = torch.arange(ymax, device=eouts.device)[None, :] < hyp_lens[:, None]
            out[:, :ymax] = torch.where(mask, trig[:, :ymax], torch.zeros_like(trig[:, :ymax]))
        return out

    def loss_fn(self, logits, ys_ctc, elens, ylens):
        """Reference op boundary (ctc.py:139-150): logits `[T, B, vocab]`, concatenated int32 targets."""
        ylens_l = [int(v) for v in ylens]
        ys, o = [], 0
        flat = ys_ctc.tolist()
        for n in ylens_l:
            ys.append(flat[o:o + n])
            o += n
        labels, ylens_d, _ = ops.pack_labels(ys, logits.device)
        elens_d = elens.to(device=logits.device, dtype=torch.int32, non_blocking=True)
        loss, _ = ctc_loss(logits.transpose(1, 0), labels, elens_d, ylens_d, self.blank, 0.0)
        return loss
