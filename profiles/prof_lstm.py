"""Time one LSTM layer recurrence (forward with saves + BPTT) on the fp32 CUDA-core kernel and on the tensor-core kernel.
Usage: python profiles/prof_lstm.py  (prints one line per shape)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_sp_b200 import ops

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for name, B, T, H, nd in [("c4 uni-LSTM 1024", 32, 250, 1024, 1), ("c1 BLSTM 256", 32, 200, 256, 2), ("BLSTM 512", 32, 250, 512, 2)]:
    gx = torch.randn(B, T, nd * 4 * H, device="cuda")
    whh = (torch.rand(nd, 4 * H, H, device="cuda") * 2 - 1) / H ** 0.5
    lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
    out = {}
    for prec in (None, "bf16"):
        y, acts, cprev, hprev = ops.lstm_seq(gx, whh, lens, nd, save=True, prec=prec)
        dy = torch.randn_like(y)
        f = timeit(lambda: ops.lstm_seq(gx, whh, lens, nd, save=True, prec=prec))
        b = timeit(lambda: ops.lstm_seq_bwd(dy, acts, cprev, whh, lens, prec=prec))
        out[prec or "fp32"] = (f, b)
    print(f"{name}: B={B} T={T}  fp32 fwd {out['fp32'][0]:.2f} ms bwd {out['fp32'][1]:.2f} ms | "
          f"tc fwd {out['bf16'][0]:.2f} ms ({out['bf16'][0] * 1e3 / T:.1f} us/step) bwd {out['bf16'][1]:.2f} ms "
          f"({out['bf16'][1] * 1e3 / T:.1f} us/step)", flush=True)
