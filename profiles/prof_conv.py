#!/usr/bin/env python3
"""Conformer conv-module core (depthwise conv + LayerNorm + Swish) forward / backward timing at the bench shapes.
NSP_CONV_PATH=legacy selects the round-1 kernels.  Algorithmic bytes: forward x + y; backward x, dy -> dz and x, dz -> dx."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_sp_b200 import ops  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


out = {}
for (B, T, d, k, dt) in [(32, 500, 512, 15, torch.bfloat16), (32, 250, 512, 15, torch.bfloat16), (32, 125, 512, 15, torch.bfloat16),
                         (32, 500, 256, 15, torch.bfloat16), (32, 125, 512, 15, torch.float32)]:
    x = torch.randn(B, T, d, device=dev).to(dt)
    dy = torch.randn(B, T, d, device=dev).to(dt)
    taps = torch.randn(k, d, device=dev) * 0.3
    bias, g, be = torch.randn(d, device=dev), torch.rand(d, device=dev) + 0.5, torch.randn(d, device=dev)
    dtaps, db, dg, dbe = torch.zeros_like(taps), torch.zeros(d, device=dev), torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    es = 2 if dt == torch.bfloat16 else 4
    ms_f = timeit(lambda: ops.conformer_conv(x, taps, bias, "layer_norm", g, be, 1e-12))
    ms_b = timeit(lambda: ops.conformer_conv_bwd(x, taps, bias, g, be, 1e-12, dy, dtaps, db, dg, dbe))
    n = B * T * d
    key = "B%d_T%d_d%d_%s" % (B, T, d, "bf16" if es == 2 else "fp32")
    out[key] = dict(fwd_ms=ms_f, fwd_gbs=2 * n * es / ms_f / 1e6, bwd_ms=ms_b, bwd_gbs=6 * n * es / ms_b / 1e6)
    print("%s: fwd %.4f ms (%.0f GB/s)  bwd %.4f ms (%.0f GB/s)" % (key, ms_f, out[key]["fwd_gbs"], ms_b, out[key]["bwd_gbs"]))
print(json.dumps({"path": os.environ.get("NSP_CONV_PATH", "stream"), "conv": out}))
