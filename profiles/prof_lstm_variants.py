"""Where the tensor-core LSTM step time goes: the C4 layer (B=32, T=250, H=1024) under the experiment knobs of lstm_tc.cu
(NSP_LSTM_TC_DEBUG 1 = no MMAs, 2 = no operand loads, 4 = no writer-side proxy fence; NSP_LSTM_TC_BWD_UPC; NSP_LSTM_TC_NST; NSP_LSTM_TC_NACC = accumulators the step's MMAs rotate over)."""
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

B, T, H, nd = 32, 250, 1024, 1
gx = torch.randn(B, T, nd * 4 * H, device="cuda")
whh = (torch.rand(nd, 4 * H, H, device="cuda") * 2 - 1) / H ** 0.5
lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
y, acts, cprev, hprev = ops.lstm_seq(gx, whh, lens, nd, save=True)
dy = torch.randn_like(y)
KEYS = ("NSP_LSTM_TC_DEBUG", "NSP_LSTM_TC_BWD_UPC", "NSP_LSTM_TC_NST", "NSP_LSTM_TC_NACC", "NSP_LSTM_TC_M")
for cfg in [{}, {"NSP_LSTM_TC_DEBUG": "1"}, {"NSP_LSTM_TC_BWD_UPC": "8"}, {"NSP_LSTM_TC_M": "128"}]:
    for k in KEYS: os.environ.pop(k, None)
    os.environ.update(cfg)
    f = timeit(lambda: ops.lstm_seq(gx, whh, lens, nd, save=True, prec="bf16"))
    b = timeit(lambda: ops.lstm_seq_bwd(dy, acts, cprev, whh, lens, prec="bf16"))
    print(f"{str(cfg):70s} fwd {f * 1e3 / T:5.2f} us/step   bwd {b * 1e3 / T:5.2f} us/step", flush=True)
