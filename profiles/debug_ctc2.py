import sys, ctypes
import numpy as np, torch
sys.path.insert(0, ".")
from neural_sp_b200 import ops
from neural_sp_b200._lib import lib, ptr, current_stream_ptr, check
from oracle import ctc_oracle
B, T, L, V = 1, 125, 112, 64
torch.manual_seed(0); rng = np.random.default_rng(0)
logits = torch.randn(B, T, V, device="cuda")
ys = [rng.integers(1, V, size=L).tolist() for _ in range(B)]
labels, ylens, _ = ops.pack_labels(ys, logits.device)
elens = torch.full((B,), T, dtype=torch.int32, device="cuda")
n = lib.nsp_ctc_loss_workspace_bytes(B, T, L)
ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
nll = torch.empty(B, device="cuda"); loss = torch.empty((), device="cuda"); grad = torch.empty(B, T, V, device="cuda")
check(lib.nsp_ctc_loss_fwd_bwd(ptr(logits), T * V, V, B, T, V, ptr(labels), L, ptr(elens), ptr(ylens), 0, 0.0, ptr(nll), ptr(loss), ptr(grad), ptr(ws), n, current_stream_ptr()))
torch.cuda.synchronize()
Sp = (2 * L + 1 + 15) // 16 * 16
lat = B * T * Sp
f = ws.view(torch.float32)
emit = f[:lat].view(T, Sp).cpu().numpy(); alpha = f[lat:2 * lat].view(T, Sp).cpu().numpy(); beta = f[2 * lat:3 * lat].view(T, Sp).cpu().numpy()
S = 2 * L + 1
lp = torch.log_softmax(logits[0].double().cpu(), -1).numpy()
path = np.zeros(S, int); path[1::2] = ys[0]
em = lp[:, path]
print("emit err", np.abs(emit[:, :S] - em).max())
# reference alpha/beta in fp64
NEG = -np.inf
la = np.full((T, S), NEG); lb = np.full((T, S), NEG)
skip = np.zeros(S, bool); skip[2:] = (path[2:] != 0) & (path[2:] != path[:-2])
la[0, 0] = em[0, 0]; la[0, 1] = em[0, 1]
for t in range(1, T):
    p0 = la[t - 1]; a1 = np.concatenate(([NEG], p0[:-1])); a2 = np.where(skip, np.concatenate(([NEG, NEG], p0[:-2])), NEG)
    la[t] = np.logaddexp(np.logaddexp(p0, a1), a2) + em[t]
lb[T - 1, S - 1] = em[T - 1, S - 1]; lb[T - 1, S - 2] = em[T - 1, S - 2]
skf = np.zeros(S, bool); skf[:-2] = skip[2:]
for t in range(T - 2, -1, -1):
    p0 = lb[t + 1]; b1 = np.concatenate((p0[1:], [NEG])); b2 = np.where(skf, np.concatenate((p0[2:], [NEG, NEG])), NEG)
    lb[t] = np.logaddexp(np.logaddexp(p0, b1), b2) + em[t]
fa = np.where(alpha[:, :S] < -1e29, NEG, alpha[:, :S]); fb = np.where(beta[:, :S] < -1e29, NEG, beta[:, :S])
for name, got, ref in (("alpha", fa, la), ("beta", fb, lb)):
    with np.errstate(invalid="ignore"):
        d = np.abs(got - ref); d[np.isnan(d)] = 0
        mism = (np.isinf(got) != np.isinf(ref))
    print(name, "max err", d[~mism].max(), "inf mismatches", mism.sum(), np.argwhere(mism)[:10].tolist())
rs = grad.sum(-1).abs()[0].cpu().numpy()
print("bad rows", np.nonzero(rs > 1e-3)[0].tolist())
nx = ws[3 * lat * 4:].cpu()
