import sys
import numpy as np, torch
sys.path.insert(0, ".")
from neural_sp_b200 import ops
for (B, T, L, V) in [(2, 250, 112, 64), (2, 250, 56, 64), (2, 125, 112, 64), (2, 40, 100, 64), (2, 300, 130, 64)]:
    torch.manual_seed(0); rng = np.random.default_rng(0)
    logits = torch.randn(B, T, V, device="cuda")
    ys = [rng.integers(1, V, size=L).tolist() for _ in range(B)]
    labels, ylens, _ = ops.pack_labels(ys, logits.device)
    elens = torch.full((B,), T, dtype=torch.int32, device="cuda")
    loss, nll, grad = ops.ctc_loss_fwd_bwd(logits, labels, elens, ylens, 0, 0.0)
    rs = grad.sum(-1).abs()
    bad = (rs > 1e-3).nonzero()
    x = logits.clone().requires_grad_(True)
    ref = torch.nn.functional.ctc_loss(x.transpose(0, 1).log_softmax(2), torch.tensor([v for y in ys for v in y], dtype=torch.int32),
                                       torch.full((B,), T, dtype=torch.int32), torch.tensor([L] * B, dtype=torch.int32), reduction="none", zero_infinity=True)
    print((B, T, L), "nll", nll.tolist(), "ref", ref.tolist(), "bad rows", bad.shape[0], bad[:12].tolist())
