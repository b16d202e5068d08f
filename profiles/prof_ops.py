"""Small drivers for ncu captures of single kernels (run under gpurun; see profiles/README.md)."""
import sys
import torch

sys.path.insert(0, ".")
from neural_sp_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "gemm_ffn1"
dev = "cuda"
torch.manual_seed(0)
if which.startswith("gemm"):
    shapes = {"gemm_ffn1": (16000, 2048, 512, "swish", True, False), "gemm_ffn2": (16000, 512, 2048, None, False, True),
              "gemm_qkv": (16000, 1536, 512, None, True, False)}
    M, N, K, act, out_bf16, res = shapes[which]
    x = torch.randn(M, K, device=dev).bfloat16()
    w = ops.prepare_weight(torch.randn(N, K, device=dev) / K ** 0.5, "bf16")
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev) if res else None
    for _ in range(5):
        y = ops.linear(x, w, b, prec="bf16", act=act, residual=r, alpha=0.5 if res else 1.0,
                       out_dtype=torch.bfloat16 if out_bf16 else torch.float32)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        y = ops.linear(x, w, b, prec="bf16", act=act, residual=r, alpha=0.5 if res else 1.0,
                       out_dtype=torch.bfloat16 if out_bf16 else torch.float32)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    print(which, "ms", ms, "TFLOP/s", 2.0 * M * N * K / ms / 1e9)
elif which == "ctc":
    B, T, V = 32, 125, 10000
    import numpy as np
    rng = np.random.default_rng(0)
    logits = torch.randn(B, T, V, device=dev)
    ys = [rng.integers(4, V, size=56).tolist() for _ in range(B)]
    labels, ylens, _ = ops.pack_labels(ys, logits.device)
    elens = torch.full((B,), T, dtype=torch.int32, device=dev)
    for (B, T, V) in [(32, 125, 10000), (32, 250, 10000), (32, 125, 1000)]:
        logits = torch.randn(B, T, V, device=dev)
        L = int(0.45 * T)
        ys = [rng.integers(4, V, size=L).tolist() for _ in range(B)]
        labels, ylens, _ = ops._pack_labels(ys, logits.device)
        elens = torch.full((B,), T, dtype=torch.int32, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for _ in range(5):
            ops.ctc_loss_fwd_bwd(logits, labels, elens, ylens, 0, 0.1)
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(20):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); ops.ctc_loss_fwd_bwd(logits, labels, elens, ylens, 0, 0.1); e.record()
            torch.cuda.synchronize(); tot += s.elapsed_time(e)
        ms = tot / 20
        print("ctc B=%d T=%d V=%d L=%d: %.4f ms/batch  %.1f GB/s algorithmic (8 B/logit)" % (B, T, V, L, ms, 8.0 * B * T * V / ms / 1e6))
elif which == "attn_bwd":
    B, T, H, dk = 32, 500, 8, 64
    D = H * dk
    qkv = (torch.randn(B, T, 3 * D, device=dev) * 0.5).bfloat16()
    r = (torch.randn(11, D, device=dev) * 0.5).bfloat16()
    klens = torch.full((B,), T, dtype=torch.int32, device=dev)
    dout = torch.randn(B, T, D, device=dev).bfloat16()
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    out, stats = ops.relpos_attention(q, k, v, klens, H, r=r, clamp_len=10, want_stats=True)
    assert stats is not None
    dr = torch.zeros(11, D, device=dev)
    for _ in range(3):
        ops.relpos_attention_bwd(q, k, v, klens, H, out, dout, r=r, clamp_len=10, dr=dr, stats=stats)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.relpos_attention_bwd(q, k, v, klens, H, out, dout, r=r, clamp_len=10, dr=dr, stats=stats)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print("attn_bwd B=%d T=%d H=%d: %.4f ms  %.1f TFLOP/s (10*T^2*d per utterance)" % (B, T, H, ms, 10.0 * B * T * T * D / ms / 1e9))
