"""CTC prefix beam search (neural_sp_b200/decoders/ctc_beam.py) hypothesis by hypothesis against the unmodified reference
(decoders/ctc.py:245-531): offline `beam_search`, the per-frame `_beam_search`, streaming `beam_search_block_sync` over several
blocks, with and without a first-pass LM (a small deterministic stand-in with the RNNLM `predict` interface on both sides).
CPU: the log-softmax kernel is replaced by its torch restatement.  Live-reference tests need /root/reference (build container);
the golden test (tests/golden/prefix_beam.npz, made by gen_golden_prefix_beam.py from the reference) runs anywhere."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.isdir("/root/reference/neural_sp")
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="reference tree not available")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

PARAMS = {'recog_beam_width': 4, 'recog_length_penalty': 0.1, 'recog_cache_embedding': False, 'recog_lm_weight': 0.3,
          'recog_lm_second_weight': 0.0, 'recog_lm_bwd_weight': 0.0, 'recog_lm_state_carry_over': False,
          'recog_softmax_smoothing': 1.0}


class ToyLM:
    """Deterministic stand-in with the interface the beam search uses (RNNLM.predict, n_layers, n_units, device)."""
    n_layers, n_units = 1, 8

    def __init__(self, vocab):
        g = torch.Generator().manual_seed(7)
        self.emb = torch.randn(vocab, 8, generator=g)
        self.out = torch.randn(8, vocab, generator=g)
        self.device = torch.device("cpu")

    def predict(self, ys, state):
        h0 = state['hxs'] if state is not None else torch.zeros(1, ys.size(0), 8)
        h = torch.tanh(self.emb[ys[:, -1]] + 0.5 * h0[0])
        scores = torch.log_softmax(h @ self.out, dim=-1).unsqueeze(1)
        return h.unsqueeze(1), {'hxs': h.unsqueeze(0), 'cxs': h.unsqueeze(0)}, scores


def _pair(vocab, fc_list, monkeypatch):
    import ops_doubles
    from oracle.ref_import import import_reference
    import_reference()
    ops_doubles.install(monkeypatch)
    from neural_sp_b200 import ops
    monkeypatch.setattr(ops, "softmax_rows", ops_doubles.softmax_rows)
    from neural_sp.models.seq2seq.decoders.ctc import CTC as RefCTC
    from neural_sp_b200.decoders.ctc import CTC
    torch.manual_seed(0)
    kw = dict(eos=2, blank=0, enc_n_units=16, vocab=vocab, dropout=0.0, lsm_prob=0.0, fc_list=fc_list)
    ref = RefCTC(**kw).eval()
    ours = CTC(**kw)
    ours.load_state_dict(ref.state_dict(), strict=True)
    return ref, ours.eval().set_precision("fp32")


def _same_hyps(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x['hyp'] == y['hyp'] and x['hyp_ids_str'] == y['hyp_ids_str'] and x['update_lm'] == y['update_lm']
        for k in ('score', 'p_b', 'p_nb', 'score_ctc', 'score_lm', 'score_lp'):
            assert abs(float(x[k]) - float(y[k])) <= 1e-9 * max(1.0, abs(float(y[k]))), (k, x[k], y[k], x['hyp'])


@needs_ref
@pytest.mark.parametrize("vocab,fc_list,beam,T", [(12, None, 4, 30), (40, "8", 5, 25), (9, None, 8, 17)])
def test_offline_beam_search_equals_reference(vocab, fc_list, beam, T, monkeypatch):
    ref, ours = _pair(vocab, fc_list, monkeypatch)
    eouts = torch.randn(3, T, 16, generator=torch.Generator().manual_seed(1)) * 2
    params = dict(PARAMS, recog_beam_width=beam, recog_lm_weight=0.0)
    with torch.no_grad():
        r = ref.beam_search(eouts, [T] * 3, params, None, nbest=min(3, beam))
    o = ours.beam_search(eouts, [T] * 3, params, None, nbest=min(3, beam))
    assert len(r) == len(o) == 3
    for rb, ob in zip(r, o):
        assert [x.tolist() for x in rb] == [x.tolist() for x in ob]


@needs_ref
@pytest.mark.parametrize("with_lm", [False, True])
@pytest.mark.parametrize("smoothing", [1.0, 0.7])
def test_block_sync_equals_reference_block_by_block(with_lm, smoothing, monkeypatch):
    from neural_sp.models.seq2seq.decoders.beam_search import BeamSearch
    vocab = 15
    ref, ours = _pair(vocab, None, monkeypatch)
    lm = ToyLM(vocab) if with_lm else None
    params = dict(PARAMS, recog_softmax_smoothing=smoothing, recog_lm_weight=0.3 if with_lm else 0.0)
    helper = BeamSearch(params['recog_beam_width'], 2, 1.0, params['recog_lm_weight'], torch.device("cpu"))
    g = torch.Generator().manual_seed(3)
    hr = ho = None
    for blk in range(4):
        eouts = torch.randn(1, 5 + blk, 16, generator=g) * 2
        with torch.no_grad():
            er, hr = ref.beam_search_block_sync(eouts, params, helper, None, hr, lm)
        eo, ho = ours.beam_search_block_sync(eouts, params, helper, None, ho, lm)
        assert er == eo == []
        _same_hyps(ho, hr)
        assert ours.n_frames == ref.n_frames
    assert list(ours.state_cache.keys()) == list(ref.state_cache.keys())


@needs_ref
def test_per_frame_search_on_given_scores(monkeypatch):
    """`_beam_search(hyps, helper, scores, lm, lp_weight)` on a fixed `[T, V]` block, including exact score ties."""
    from neural_sp.models.seq2seq.decoders.beam_search import BeamSearch
    from collections import OrderedDict
    ref, ours = _pair(11, None, monkeypatch)
    g = torch.Generator().manual_seed(5)
    scores = torch.log_softmax(torch.randn(20, 11, generator=g) * 3, dim=-1)
    scores[4] = scores[3]                                   # repeated frame
    scores[7, 1:] = scores[7, 1:].mean()                    # a frame whose tokens all tie
    scores[7] = torch.log_softmax(scores[7], dim=-1)
    helper = BeamSearch(3, 2, 1.0, 0.0, torch.device("cpu"))
    ref.state_cache, ours.state_cache = OrderedDict(), OrderedDict()
    hr, nr = ref._beam_search(ref.initialize_beam([2], None), helper, scores, None, 0.2)
    ho, no = ours._beam_search(ours.initialize_beam([2], None), helper, scores, None, 0.2)
    _same_hyps(ho, hr)
    _same_hyps(no, nr)


def test_golden_beams():
    """Fixtures from the reference (gen_golden_prefix_beam.py): frame scores in, hypotheses out; runs without the reference."""
    from neural_sp_b200.decoders import ctc_beam
    from collections import OrderedDict
    g = np.load(os.path.join(GOLDEN, "prefix_beam.npz"), allow_pickle=True)
    for c in range(int(g["n_cases"])):
        block = g["scores_%d" % c].astype(np.float64)
        beam, lp = int(g["beam_%d" % c]), float(g["lp_%d" % c])
        hyps, _ = ctc_beam.prefix_beam_search(ctc_beam.initialize_beam([2], None), block, beam, block.shape[1], 0, None, 0.0, lp,
                                              OrderedDict())
        want = g["hyps_%d" % c]
        assert [h['hyp'] for h in hyps] == [list(w) for w in want]
        np.testing.assert_allclose([h['score'] for h in hyps], g["scores_out_%d" % c], rtol=1e-9, atol=1e-9)


@needs_ref
def test_per_frame_search_many_random_blocks(monkeypatch):
    """60 random `[T, V]` blocks (beam 1-9, vocabularies down to beam + 1, peaked and flat distributions, quantised scores that
    tie often, length penalties of both signs) against the reference's `_beam_search`."""
    from collections import OrderedDict
    from oracle.ref_import import import_reference
    import_reference()
    from neural_sp.models.seq2seq.decoders.beam_search import BeamSearch
    rng = np.random.RandomState(0)
    for case in range(60):
        beam = int(rng.randint(1, 10))
        V = int(rng.randint(beam + 2, 40))
        T = int(rng.randint(1, 35))
        ref, ours = _pair(V, None, monkeypatch)
        g = torch.Generator().manual_seed(case)
        raw = torch.randn(T, V, generator=g) * float(rng.choice([0.3, 1.0, 4.0]))
        if case % 3 == 0:
            raw = (raw * 2).round() / 2                      # many exact ties between tokens
        scores = torch.log_softmax(raw, dim=-1)
        lp = float(rng.choice([0.0, 0.2, -0.1]))
        helper = BeamSearch(beam, 2, 1.0, 0.0, torch.device("cpu"))
        ref.state_cache, ours.state_cache = OrderedDict(), OrderedDict()
        hr, nr = ref._beam_search(ref.initialize_beam([2], None), helper, scores, None, lp)
        ho, no = ours._beam_search(ours.initialize_beam([2], None), helper, scores, None, lp)
        try:
            _same_hyps(ho, hr)
            _same_hyps(no, nr)
        except AssertionError as e:
            raise AssertionError("case %d (beam %d, V %d, T %d, lp %g): %s" % (case, beam, V, T, lp, e))
