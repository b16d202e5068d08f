"""Golden CTC prefix-beam-search results from the UNMODIFIED reference (decoders/ctc.py:365-483 `_beam_search`), CPU.
Run in the build container:  python tests/golden/gen_golden_prefix_beam.py   -> tests/golden/prefix_beam.npz"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference   # noqa: E402

import_reference()
from neural_sp.models.seq2seq.decoders.beam_search import BeamSearch   # noqa: E402
from neural_sp.models.seq2seq.decoders.ctc import CTC   # noqa: E402

out = {}
cases = [(25, 12, 4, 0.0, 1), (40, 30, 6, 0.15, 2), (18, 8, 3, 0.3, 3), (60, 50, 10, 0.05, 4)]
for c, (T, V, beam, lp, seed) in enumerate(cases):
    g = torch.Generator().manual_seed(seed)
    scores = torch.log_softmax(torch.randn(T, V, generator=g) * 3, dim=-1)
    if c == 2:
        scores[5] = scores[4]
    ctc = CTC(eos=2, blank=0, enc_n_units=4, vocab=V)
    ctc.state_cache = OrderedDict()
    helper = BeamSearch(beam, 2, 1.0, 0.0, torch.device("cpu"))
    hyps, _ = ctc._beam_search(ctc.initialize_beam([2], None), helper, scores, None, lp)
    out["scores_%d" % c] = scores.numpy()
    out["beam_%d" % c] = beam
    out["lp_%d" % c] = lp
    hl = np.empty(len(hyps), dtype=object)
    for i, h in enumerate(hyps):
        hl[i] = np.array(h['hyp'], dtype=np.int64)
    out["hyps_%d" % c] = hl
    out["scores_out_%d" % c] = np.array([float(h['score']) for h in hyps])
out["n_cases"] = len(cases)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "prefix_beam.npz"), **out)
print("wrote prefix_beam.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.startswith("hyps")})
