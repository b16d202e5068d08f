#!/usr/bin/env python3
"""CTC fwd+bwd timing at the BASELINE sizes: CUDA events around the whole C-ABI call (memset + kernels), L2 flushed between
iterations, algorithmic bytes = 8 B per logit.  NSP_CTC_PATH=legacy selects the round-1 three-kernel path."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_sp_b200 import ops  # noqa: E402

peak = 6582.5
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for (B, T, V) in [(32, 125, 10000), (32, 250, 10000), (32, 125, 1000), (2, 200, 32)]:
    rng = np.random.default_rng(0)
    logits = torch.randn(B, T, V, device=dev)
    L = max(1, int(0.45 * T))
    ys = [rng.integers(4, V, size=L).tolist() for _ in range(B)]
    labels, ylens, _ = ops.pack_labels(ys, dev)
    elens = torch.full((B,), T, dtype=torch.int32, device=dev)
    for _ in range(5):
        ops.ctc_loss_fwd_bwd(logits, labels, elens, ylens, 0, 0.1)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.ctc_loss_fwd_bwd(logits, labels, elens, ylens, 0, 0.1)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    gbs = 8.0 * B * T * V / (ms * 1e-3) / 1e9
    out["B%d_T%d_V%d" % (B, T, V)] = dict(ms=ms, min_ms=float(min(ts)), gbs=gbs, frac=gbs / peak)
    print("ctc B=%d T'=%d V=%d: %.4f ms (min %.4f)  %.0f GB/s algorithmic = %.1f %% of %.0f" % (B, T, V, ms, min(ts), gbs, 100 * gbs / peak, peak))
    if int(os.environ.get("NSP_CTC_DEBUG", "0")) & 8 and ops.LAST_CTC_WS is not None:
        torch.cuda.synchronize()
        n = 2 * 160 + 4 * B + 64
        tr = ops.LAST_CTC_WS[-(n * 8 + 256):-256].view(torch.int64).cpu().numpy().astype(np.int64)
        grid = min(148, B * T if V > 1280 else 10 ** 9)
        cta = tr[:2 * 160].reshape(160, 2)
        live = cta[:, 0] > 0
        t0 = cta[live, 0].min()
        print("   trace: %d CTAs, start spread %.1f us, rows done at %.1f .. %.1f us" % (live.sum(), (cta[live, 0].max() - t0) / 1e3,
              (cta[live, 1].min() - t0) / 1e3, (cta[live, 1].max() - t0) / 1e3))
        sw = tr[2 * 160:2 * 160 + 4 * B].reshape(B, 2, 2)
        rows = tr[2 * 160 + 4 * B:].reshape(16, 4)
        print('   CTA 40 rows: wait %s | compute %s | fence+arrive %s | row period %s (us)' % (
            ' '.join('%.2f' % ((r[1] - r[0]) / 1e3) for r in rows if r[0] > 0), ' '.join('%.2f' % ((r[2] - r[1]) / 1e3) for r in rows if r[0] > 0),
            ' '.join('%.2f' % ((r[3] - r[2]) / 1e3) for r in rows if r[0] > 0),
            ' '.join('%.2f' % ((rows[i + 1][0] - rows[i][0]) / 1e3) for i in range(15) if rows[i + 1][0] > 0)))
        print('   alpha us/step:', ' '.join('%.2f' % ((sw[b, 0, 1] - sw[b, 0, 0]) / 1e3 / T) for b in range(B)))
        print('   beta  us/step:', ' '.join('%.2f' % ((sw[b, 1, 1] - sw[b, 1, 0]) / 1e3 / T) for b in range(B)))
        print('   alpha start  :', ' '.join('%.0f' % ((sw[b, 0, 0] - t0) / 1e3) for b in range(B)))
        print('   alpha end    :', ' '.join('%.0f' % ((sw[b, 0, 1] - t0) / 1e3) for b in range(B)))
        print('   beta end     :', ' '.join('%.0f' % ((sw[b, 1, 1] - t0) / 1e3) for b in range(B)))
        for b in sorted(set([0, B - 1])):
            if 0 <= b < B:
                print("   utt %2d: alpha %.1f -> %.1f us (%.3f us/step)   beta %.1f -> %.1f us" % (
                    b, (sw[b, 0, 0] - t0) / 1e3, (sw[b, 0, 1] - t0) / 1e3, (sw[b, 0, 1] - sw[b, 0, 0]) / 1e3 / T,
                    (sw[b, 1, 0] - t0) / 1e3, (sw[b, 1, 1] - t0) / 1e3))
print(json.dumps({"path": os.environ.get("NSP_CTC_PATH", "stream"), "ctc": out}))
