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
elif which == "wgrad":
    M, N, K = 16000, 2048, 512            # dW [N, K] += dY^T X,  dY [M, N], X [M, K]
    dy = torch.randn(M, N, device=dev).bfloat16()
    x = torch.randn(M, K, device=dev).bfloat16()
    dw = torch.zeros(N, K, device=dev)
    for _ in range(5):
        ops.linear_wgrad(dy, x, "bf16", dw)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        ops.linear_wgrad(dy, x, "bf16", dw)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    print(which, "ms", ms, "TFLOP/s", 2.0 * M * N * K / ms / 1e9)
elif which == "attn_fwd":
    B, T, H, dk = 32, 500, 8, 64
    D = H * dk
    qkv = (torch.randn(B, T, 3 * D, device=dev) * 0.5).bfloat16()
    r = (torch.randn(11, D, device=dev) * 0.5).bfloat16()
    klens = torch.full((B,), T, dtype=torch.int32, device=dev)
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    for _ in range(3):
        ops.relpos_attention(q, k, v, klens, H, r=r, clamp_len=10)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.relpos_attention(q, k, v, klens, H, r=r, clamp_len=10)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print("attn_fwd B=%d T=%d H=%d: %.4f ms  %.1f TFLOP/s (4*T^2*d per utterance)" % (B, T, H, ms, 4.0 * B * T * T * D / ms / 1e9))
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
elif which == "stream_kernels":
    # HBM-bound kernels added late in round 1: time each on a tensor much larger than L2, report algorithmic GB/s
    from neural_sp_b200 import random as nrandom

    def timeit(name, fn, nbytes, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / reps
        print("%-34s %8.4f ms  %8.1f GB/s algorithmic" % (name, ms, nbytes / ms / 1e6))

    M, dff, d = 16000, 2048, 512
    hb = torch.randn(M, dff, device=dev).bfloat16()
    sid = nrandom.next_stream()
    timeit("dropout bf16 [16000,2048] in place", lambda: ops.dropout(hb, 0.1, sid, inplace=True), hb.numel() * 4)
    t, res = torch.randn(M, d, device=dev).bfloat16(), torch.randn(M, d, device=dev)
    out = torch.empty_like(res)
    timeit("dropout_add bf16->fp32 [16000,512]", lambda: ops.dropout_add(t, res, 0.1, 0.5, sid, out=out), M * d * 10)
    B, T, U1, V = 32, 250, 57, 1000
    lp = torch.log_softmax(torch.randn(B, T, U1, V, device=dev), -1)
    ys = torch.randint(1, V, (B, U1 - 1), dtype=torch.int32, device=dev)
    fl = torch.full((B,), T, dtype=torch.int32, device=dev)
    yl = torch.full((B,), U1 - 1, dtype=torch.int32, device=dev)
    loss, nll, _, ws = ops.rnnt_loss_fwd_bwd(lp, ys, fl, yl, 0, need_grad=False, return_ws=True)
    timeit("rnnt lattice (gather+alpha/beta)", lambda: ops.rnnt_loss_fwd_bwd(lp, ys, fl, yl, 0, need_grad=False), B * T * U1 * 8.0)
    timeit("rnnt_grad_logits -> bf16", lambda: ops.rnnt_grad_logits(lp, ws, nll, ys, fl, yl, 0, out_dtype=torch.bfloat16), lp.numel() * 6.0, 5)
    timeit("rnnt dense grad (old path)", lambda: ops.rnnt_loss_fwd_bwd(lp, ys, fl, yl, 0, need_grad=True), lp.numel() * 4.0, 5)
    J = 640
    h = torch.tanh(torch.randn(8, T, U1, J, device=dev)).bfloat16()
    dh = torch.randn_like(h)
    timeit("rnnt_joint_tanh_bwd bf16 [8,250,57,640]", lambda: ops.rnnt_joint_tanh_bwd(h, dh), h.numel() * 2 * 2 * 2.0, 5)
    g = torch.randn(32, 500, 512, device=dev).bfloat16()
    taps, bias = torch.randn(15, 512, device=dev) * 0.2, torch.zeros(512, device=dev)
    timeit("dwconv_stats bf16 [32,500,512] k15", lambda: ops.dwconv_stats(g, taps, bias), g.numel() * 4.0)
    z, stats = ops.dwconv_stats(g, taps, bias)
    mean = (stats[0] / (32 * 500)).contiguous()
    var = (stats[1] / (32 * 500) - mean * mean).clamp_min(0).contiguous()
    gam, bet = torch.ones(512, device=dev), torch.zeros(512, device=dev)
    timeit("bn_swish_bwd bf16 [32,500,512]", lambda: ops.bn_swish_bwd(z, g, mean, var, gam, bet, 1e-5), g.numel() * 10.0)
    x3 = torch.randn(32, 500, 512, device=dev)
    dy3 = torch.randn(32, 250, 512, device=dev)
    timeit("pool_time_bwd mean fp32 [32,500,512]", lambda: ops.pool_time_bwd(dy3, 500, 2, "mean"), x3.numel() * 6.0)
