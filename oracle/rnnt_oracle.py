"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the RNN-Transducer loss at the reference's op boundary.

Reference call site: neural_sp/models/seq2seq/decoders/rnn_transducer.py:248-252
    warp_rnnt.rnnt_loss(log_probs, ys_out, elens, ylens, average_frames=False, reduction='mean', gather=False)
(CPU twin warprnnt_pytorch.RNNTLoss(), :254-256).  The arithmetic lives in third-party packages that are
NOT in the reference tree and cannot be installed offline: warp_rnnt==0.3 (tools/Makefile:144-146) and
HawkAaron/warp-transducer at un-pinned HEAD (tools/Makefile:133-142).  This file restates the published
algorithm (Graves 2012, "Sequence Transduction with Recurrent Neural Networks", eq. 16-20) with blank = 0,
mean over the batch and no frame averaging.

PARITY UNPINNED by the reference: its only test of this path (test/decoders/test_rnn_transducer_decoder.py:62-84)
asserts loss >= 0 and shape.  The restatement is pinned instead against torchaudio.functional.rnnt_loss
(tests/golden/rnnt_*.npz), an independent implementation of the same published algorithm.
"""
import numpy as np


def rnnt_nll_and_grad(log_probs, ys, flens, ylens, blank=0):
    """log_probs [B,T,U+1,V] -> (nll [B], loss = mean nll, d loss / d log_probs [B,T,U+1,V])."""
    lp = np.asarray(log_probs, dtype=np.float64)
    B, T, U1, V = lp.shape
    nll = np.zeros(B)
    grad = np.zeros_like(lp)
    for b in range(B):
        Tb, Ub = int(flens[b]), int(ylens[b])
        y = [int(v) for v in ys[b][:Ub]]
        a = np.full((Tb, Ub + 1), -np.inf)
        be = np.full((Tb, Ub + 1), -np.inf)
        a[0, 0] = 0.0
        for t in range(Tb):
            for u in range(Ub + 1):
                if t == 0 and u == 0:
                    continue
                v1 = a[t - 1, u] + lp[b, t - 1, u, blank] if t > 0 else -np.inf
                v2 = a[t, u - 1] + lp[b, t, u - 1, y[u - 1]] if u > 0 else -np.inf
                a[t, u] = np.logaddexp(v1, v2)
        be[Tb - 1, Ub] = lp[b, Tb - 1, Ub, blank]
        for t in range(Tb - 1, -1, -1):
            for u in range(Ub, -1, -1):
                if t == Tb - 1 and u == Ub:
                    continue
                v1 = be[t + 1, u] + lp[b, t, u, blank] if t < Tb - 1 else -np.inf
                v2 = be[t, u + 1] + lp[b, t, u, y[u]] if u < Ub else -np.inf
                be[t, u] = np.logaddexp(v1, v2)
        nll[b] = -be[0, 0]
        for t in range(Tb):
            for u in range(Ub + 1):
                if t < Tb - 1:
                    grad[b, t, u, blank] -= np.exp(a[t, u] + lp[b, t, u, blank] + be[t + 1, u] + nll[b])
                elif u == Ub:
                    grad[b, t, u, blank] -= np.exp(a[t, u] + lp[b, t, u, blank] + nll[b])
                if u < Ub:
                    grad[b, t, u, y[u]] -= np.exp(a[t, u] + lp[b, t, u, y[u]] + be[t, u + 1] + nll[b])
    return nll, nll.mean(), grad / B
