"""CTC prefix beam search, offline and block-synchronous (streaming): the host half of
reference decoders/ctc.py:245-531 (`initialize_beam`, `beam_search`, `_beam_search`, `beam_search_block_sync`).

What runs where.  The device produces the frame scores (output head + log-softmax kernel, `CTC.scores`) and ONE copy brings
the `[T, V]` block to pinned host memory; everything after that is per-frame bookkeeping over at most
`beam_width * (beam_width + 1)` candidates.  The reference does that bookkeeping with one `torch.topk` launch per frame and one
`.item()` device synchronisation per (frame, beam, candidate); here the top-k of all frames is ONE batched `torch.topk` on
the host block and the candidate scores of a frame are formed as arrays (one row per live hypothesis, one column per
extension), ranked with a stable sort and merged by token sequence.

Same observable behaviour as the reference, which the tests pin hypothesis by hypothesis (tests/test_ctc_beam_cpu.py against
the live reference, tests/golden/ctc_beam_*.npz elsewhere):
  * candidate order = hypothesis-major, "not extended" first, then the top-k tokens in descending score order; ranking is a
    stable descending sort on the total score, so ties resolve as in the reference's `sorted(..., reverse=True)`;
  * hypotheses with the same token sequence are merged in rank order: the first keeps its entry, later ones add their
    `score` and `score_ctc` to it in the log domain (its `p_b` / `p_nb` stay: the reference's `merge_ctc_path(merge_prob=True)`);
  * scores are float64 sums of the float32 frame scores (the reference adds Python floats);
  * the hypothesis records are the reference's dicts (`hyp`, `hyp_ids_str`, `score`, `p_b`, `p_nb`, `score_ctc`, `score_lm`,
    `score_lp`, `next_scores_lm`, `lmstate`, `update_lm`): `beam_search_block_sync` hands them back to its caller
    (speech2text.py:627), which passes them in again with the next block.
Shallow fusion with a first-pass LM follows the reference's protocol (batched `lm.predict(ys, state)` for the hypotheses whose
last token is new, results cached by token sequence); the LM itself is outside this package (duck-typed).
"""
from collections import OrderedDict

import numpy as np
import torch

LOG_0 = -1e10          # reference ctc.py:28-29
LOG_1 = 0


def initialize_beam(hyp, lmstate):
    """One empty hypothesis (reference ctc.py:245-254)."""
    return [{'hyp': hyp, 'hyp_ids_str': '', 'p_b': LOG_1, 'p_nb': LOG_0, 'score_lm': LOG_1, 'lmstate': lmstate,
             'update_lm': True}]


def frame_scores_to_host(log_probs):
    """`[T, V]` device tensor -> fp32 host tensor (one copy through pinned memory)."""
    lp = log_probs.detach().float()
    if lp.is_cuda:
        host = torch.empty(lp.shape, dtype=torch.float32, pin_memory=True)
        host.copy_(lp, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        lp = host
    return lp.contiguous()


def _topk_tokens(block32, k):
    """Per frame, the k best tokens of columns 1.. (column 0 = blank is excluded like the reference's `scores_ctc[t, 1:]`) in
    descending score order: ONE batched `torch.topk` over the host block.  Exact ties between fp32 scores resolve as torch's
    CPU top-k resolves them, which is what the reference does when it runs on the host (on a GPU its order among tied tokens
    is whatever the CUDA top-k returns)."""
    if block32.shape[0] == 0:
        return np.zeros((0, k), dtype=np.int64)
    _, ids = torch.topk(block32[:, 1:], k=k, dim=-1, largest=True, sorted=True)
    return ids.numpy().astype(np.int64) + 1


def _update_lm(hyps, lm, state_cache):
    """Advance the LM for the hypotheses whose last token has not been scored yet (reference ctc.py:377-409), in one batch."""
    todo = [h for h in hyps if h['update_lm']]
    if not todo:
        return
    scores_lm, states = None, None
    if lm is not None:
        ys = torch.zeros((len(todo), 1), dtype=torch.int64, device=getattr(lm, 'device', 'cpu'))
        for i, h in enumerate(todo):
            ys[i] = h['hyp'][-1]
        prev = None
        if todo[0]['lmstate'] is not None:
            prev = {'hxs': torch.cat([h['lmstate']['hxs'] for h in todo], dim=1),
                    'cxs': torch.cat([h['lmstate']['cxs'] for h in todo], dim=1)}
        _, states, scores_lm = lm.predict(ys, prev)
    for i, h in enumerate(todo):
        h['lmstate'] = {'hxs': states['hxs'][:, i:i + 1], 'cxs': states['cxs'][:, i:i + 1]} if states is not None else None
        h['next_scores_lm'] = scores_lm[i:i + 1] if lm is not None else None
        h['update_lm'] = False
        state_cache[h['hyp_ids_str']] = {'next_scores_lm': h['next_scores_lm'], 'lmstate': h['lmstate']}


def prefix_beam_search(hyps, block, beam_width, vocab, blank, lm, lm_weight, lp_weight, state_cache):
    """Advance `hyps` over the frames of `block` (`[T, V]` fp32 log-probabilities, host tensor or array).
    -> (hyps, candidates of the last frame).  The arithmetic of reference ctc.py:365-483, in float64."""
    block32 = torch.as_tensor(np.asarray(block, dtype=np.float32)) if not torch.is_tensor(block) else block.float()
    block = block32.numpy().astype(np.float64)
    T = block.shape[0]
    lm_weight = 0.0 if lm_weight is None else lm_weight
    lp_weight = 0.0 if lp_weight is None else lp_weight
    k = min(beam_width, vocab, max(block.shape[1] - 1, 1))
    topk = _topk_tokens(block32, k) if T > 0 else None
    # the reference indexes topk_ids[kk] for kk < beam_width: with vocab - 1 < beam_width it would fail; same guard here
    n_ext = min(beam_width, topk.shape[1]) if T > 0 else 0
    new_hyps = []
    for t in range(T):
        _update_lm(hyps, lm, state_cache)
        row = block[t]
        ids = topk[t, :n_ext]
        p_tok = row[ids]                                           # [k]
        n = len(hyps)
        p_b = np.array([h['p_b'] for h in hyps], dtype=np.float64)
        p_nb = np.array([h['p_nb'] for h in hyps], dtype=np.float64)
        s_lm = np.array([h['score_lm'] for h in hyps], dtype=np.float64)
        n_tok = np.array([len(h['hyp']) - 1 for h in hyps], dtype=np.int64)
        last = np.array([h['hyp'][-1] if len(h['hyp']) > 1 else -1 for h in hyps], dtype=np.int64)

        # column 0: the hypothesis is not extended
        stay_b = np.logaddexp(p_b + row[blank], p_nb + row[blank])
        stay_nb = np.where(n_tok > 0, p_nb + row[np.maximum(last, 0)], LOG_0)
        stay_ctc = np.logaddexp(stay_b, stay_nb)
        stay_lp = n_tok * lp_weight
        stay_total = stay_ctc + stay_lp + s_lm * lm_weight

        # columns 1..k: extended by token ids[j]; a repeated last token can only follow a blank
        same = last[:, None] == ids[None, :]
        ext_nb = np.where(same, (p_b[:, None] + p_tok[None, :]),
                          np.logaddexp(p_b[:, None] + p_tok[None, :], p_nb[:, None] + p_tok[None, :]))
        ext_ctc = np.logaddexp(LOG_0, ext_nb)
        ext_lp = (n_tok[:, None] + 1) * lp_weight
        ext_lm = np.repeat(s_lm[:, None], n_ext, axis=1)
        if lm is not None:
            # reference ctc.py:449-450 adds the LM score of candidate k to a running total that is NOT reset between the
            # candidates of one hypothesis (`total_score_lm += ...` inside the k loop): candidate k carries the LM scores of
            # candidates 0..k.  Reproduced as is: this module's contract is the reference's output.
            for i, h in enumerate(hyps):
                nxt = h['next_scores_lm'][0, 0]
                vals = nxt[torch.as_tensor(ids, device=nxt.device)].double().cpu().numpy() if torch.is_tensor(nxt) else np.asarray(nxt, dtype=np.float64)[ids]
                ext_lm[i] += np.cumsum(vals)
        ext_total = ext_ctc + ext_lp
        ext_total = ext_total + ext_lm * lm_weight

        # rank: hypothesis-major, stay first, then the tokens in top-k order; stable descending sort
        total = np.concatenate([stay_total[:, None], ext_total], axis=1).reshape(-1)
        rank = np.argsort(-total, kind='stable')

        merged = OrderedDict()
        for r in rank:
            i, j = divmod(int(r), n_ext + 1)
            h = hyps[i]
            if j == 0:
                key = h['hyp_ids_str']
                if key in merged:
                    m = merged[key]
                    m['score'] = np.logaddexp(m['score'], stay_total[i])
                    m['score_ctc'] = np.logaddexp(m['score_ctc'], stay_ctc[i])
                    continue
                merged[key] = {'hyp': h['hyp'][:], 'hyp_ids_str': key, 'score': stay_total[i], 'p_b': stay_b[i],
                               'p_nb': stay_nb[i] if n_tok[i] > 0 else LOG_0, 'score_ctc': stay_ctc[i], 'score_lm': h['score_lm'],
                               'score_lp': stay_lp[i], 'next_scores_lm': h['next_scores_lm'], 'lmstate': h['lmstate'],
                               'update_lm': False}
            else:
                tok = int(ids[j - 1])
                seq = h['hyp'] + [tok]
                key = ' '.join(map(str, seq))
                if key in merged:
                    m = merged[key]
                    m['score'] = np.logaddexp(m['score'], ext_total[i, j - 1])
                    m['score_ctc'] = np.logaddexp(m['score_ctc'], ext_ctc[i, j - 1])
                    continue
                cached = state_cache.get(key)
                merged[key] = {'hyp': seq, 'hyp_ids_str': key, 'score': ext_total[i, j - 1], 'p_b': LOG_0,
                               'p_nb': ext_nb[i, j - 1], 'score_ctc': ext_ctc[i, j - 1], 'score_lm': ext_lm[i, j - 1],
                               'score_lp': ext_lp[i, 0],
                               'next_scores_lm': cached['next_scores_lm'] if cached is not None else None,
                               'lmstate': cached['lmstate'] if cached is not None else h['lmstate'],
                               'update_lm': cached is None}
        new_hyps = list(merged.values())
        hyps = new_hyps[:beam_width]
    return hyps, new_hyps
