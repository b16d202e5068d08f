"""LayerNorm, rel-pos attention and the Conformer conv core against plain PyTorch fp32 references of the
same ops (written from the reference's formulas: relative_multihead_attention.py:146-220,
conformer_convolution.py:113-124).  fp32 paths: 2e-5 of max|ref|; bf16 I/O paths: 2e-2."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,D", [(1000, 256), (333, 512), (50, 8), (77, 80), (64, 1024), (9, 100)])
def test_layernorm(M, D):
    from neural_sp_b200 import ops
    torch.manual_seed(D)
    x = torch.randn(M, D, device="cuda") * 3 + 1
    w, b = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    ref = F.layer_norm(x.double(), (D,), w.double(), b.double(), 1e-12)
    y, yb = ops.layernorm(x, w, b, 1e-12, out_fp32=True, out_bf16=True)
    assert (y.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    assert (yb.double() - ref).abs().max().item() <= 8e-3 * ref.abs().max().item()


def _attn_ref(q, k, v, r, u, vb, klens, H, clamp, causal, lookahead, chunk_c, chunk_l):
    B, Tq, D = q.shape
    Tk = k.shape[1]
    dk = D // H
    mlen = Tk - Tq
    q4, k4, v4 = (t.double().view(B, -1, H, dk) for t in (q, k, v))
    qu = q4 + (u.double()[None, None] if u is not None else 0)
    qv = q4 + (vb.double()[None, None] if vb is not None else 0)
    e = torch.einsum("bihd,bjhd->bijh", qu, k4)
    if r is not None:
        r3 = r.double().view(-1, H, dk)
        bd_raw = torch.einsum("bihd,rhd->birh", qv, r3)
        i = torch.arange(Tq, device=q.device)[:, None]
        j = torch.arange(Tk, device=q.device)[None, :]
        dist = (mlen + i - j).abs()
        if clamp > 0:
            dist = dist.clamp(max=clamp)
        dist = dist.clamp(max=r3.shape[0] - 1)
        bd = torch.gather(bd_raw, 2, dist[None, :, :, None].expand(B, Tq, Tk, H))
        e = e + bd
    e = e / math.sqrt(dk)
    i = torch.arange(Tq, device=q.device)[None, :, None]
    j = torch.arange(Tk, device=q.device)[None, None, :]
    mask = j < klens[:, None, None]
    if causal:
        mask = mask & (j <= mlen + i + lookahead)
    if chunk_c > 0:
        cs = ((mlen + i) // chunk_c) * chunk_c
        mask = mask & (j >= cs - chunk_l) & (j < cs + chunk_c)
    e = e.masked_fill(~mask[..., None], torch.finfo(torch.float32).min)
    aw = torch.softmax(e, dim=2)
    return torch.einsum("bijh,bjhd->bihd", aw, v4).reshape(B, Tq, D)


@pytest.mark.parametrize("cfg", [
    dict(B=3, T=125, H=4, dk=64, clamp=10), dict(B=2, T=250, H=8, dk=64, clamp=10),
    dict(B=2, T=97, H=4, dk=64, clamp=-1), dict(B=2, T=40, H=4, dk=2, clamp=-1),
    dict(B=2, T=70, H=2, dk=128, clamp=5, xl=True), dict(B=2, T=64, H=4, dk=16, clamp=-1, rel=False),
    dict(B=2, T=90, H=4, dk=64, clamp=10, causal=True, lookahead=2),
    dict(B=2, T=96, H=4, dk=64, clamp=-1, chunk_c=16, chunk_l=32),
    dict(B=2, T=33, Tk=80, H=4, dk=64, clamp=10),
    dict(B=2, T=300, H=2, dk=64, clamp=10, chunk_c=32, chunk_l=64), dict(B=2, T=200, H=4, dk=64, clamp=-1, rel=False),
    dict(B=3, T=500, H=8, dk=64, clamp=10), dict(B=2, T=129, H=1, dk=64, clamp=1),
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_relpos_attention(cfg, dtype):
    from neural_sp_b200 import ops
    B, T, H, dk = cfg["B"], cfg["T"], cfg["H"], cfg["dk"]
    Tk = cfg.get("Tk", T)
    D = H * dk
    torch.manual_seed(T + dk)
    dev = "cuda"
    qkv = torch.randn(B, Tk, 3 * D, device=dev).to(dtype)
    q, k, v = qkv[:, Tk - T:, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    q = q.contiguous() if Tk != T else q
    r = (torch.randn(Tk, D, device=dev).to(dtype)) if cfg.get("rel", True) else None
    u = torch.randn(H, dk, device=dev) * 0.3 if cfg.get("xl") else None
    vb = torch.randn(H, dk, device=dev) * 0.3 if cfg.get("xl") else None
    klens = torch.tensor([Tk - 7 * b for b in range(B)], dtype=torch.int32, device=dev)
    ref = _attn_ref(q.float(), k.float(), v.float(), r.float() if r is not None else None, u, vb, klens, H,
                    cfg["clamp"], cfg.get("causal", False), cfg.get("lookahead", 0), cfg.get("chunk_c", 0), cfg.get("chunk_l", 0))
    out = ops.relpos_attention(q, k, v, klens, H, r=r, u_bias=u, v_bias=vb, clamp_len=cfg["clamp"],
                               causal=cfg.get("causal", False), lookahead=cfg.get("lookahead", 0),
                               chunk_c=cfg.get("chunk_c", 0), chunk_l=cfg.get("chunk_l", 0))
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= tol, err


@pytest.mark.parametrize("d,k,T,mode,causal", [(256, 15, 125, "layer_norm", False), (512, 15, 250, "layer_norm", False),
                                                (512, 31, 77, "layer_norm", False), (8, 3, 45, "batch_norm", False),
                                                (64, 7, 50, "layer_norm", True), (64, 5, 33, "group_norm", False),
                                                (144, 9, 40, "batch_norm", True)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conformer_conv(d, k, T, mode, causal, dtype):
    from neural_sp_b200 import ops
    torch.manual_seed(d + k)
    dev = "cuda"
    B = 3
    x = torch.randn(B, T, d, device=dev).to(dtype)
    w = torch.randn(d, 1, k, device=dev) * 0.3
    bias = torch.randn(d, device=dev) * 0.1
    g, bt = torch.rand(d, device=dev) + 0.5, torch.randn(d, device=dev) * 0.1
    rm, rv = torch.randn(d, device=dev) * 0.1, torch.rand(d, device=dev) + 0.5
    pad = k - 1 if causal else (k - 1) // 2
    xr = x.float().double().transpose(1, 2)
    y = F.conv1d(xr, w.double(), bias.double(), padding=pad, groups=d)
    if causal:
        y = y[:, :, :-pad]
    y = y.transpose(1, 2)
    if mode == "layer_norm":
        eps = 1e-12
        y = F.layer_norm(y, (d,), g.double(), bt.double(), eps)
    elif mode == "batch_norm":
        eps = 1e-5
        y = (y - rm.double()) / torch.sqrt(rv.double() + eps) * g.double() + bt.double()
    else:
        eps = 1e-5
        y = F.group_norm(y.reshape(B * T, d, 1), d // 2, g.double(), bt.double(), eps).reshape(B, T, d)
    ref = y * torch.sigmoid(y)
    out = ops.conformer_conv(x, w, bias, mode, g, bt, eps, rm, rv, causal=causal)
    tol = 3e-5 if dtype == torch.float32 else 2e-2
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= tol, err


@pytest.mark.parametrize("B,T,F,pool", [(2, 64, 80, False), (3, 50, 80, True), (2, 37, 40, True), (1, 8, 16, False),
                                        (2, 33, 21, True)])
def test_conv3x3_tc_matches_fp32_conv(B, T, F, pool):
    """tcgen05 implicit-GEMM conv vs F.conv2d (+max_pool2d ceil) on the same bf16-rounded operands."""
    from neural_sp_b200 import ops
    torch.manual_seed(T)
    dev = "cuda"
    x = torch.randn(B, T, F, 32, device=dev).bfloat16()
    w = (torch.randn(32, 32, 3, 3, device=dev) / 17.0)
    bias = torch.randn(32, device=dev) * 0.1
    wt = w.permute(0, 2, 3, 1).reshape(32, 288).bfloat16().contiguous()
    y = ops.conv3x3_c32_tc(x, wt, bias, relu=True, pool2x2=pool)
    ref = F_conv_ref(x.float(), w.bfloat16().float(), bias, pool)
    assert y.shape == ref.shape
    err = (y.float() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 1e-2, err


def F_conv_ref(x_cl, w, bias, pool):
    y = torch.relu(F.conv2d(x_cl.permute(0, 3, 1, 2).double(), w.double(), bias.double(), padding=1))
    if pool:
        y = F.max_pool2d(y, 2, 2, 0, ceil_mode=True)
    return y.permute(0, 2, 3, 1).float()


def test_frontend_simt_conv_pool_layouts():
    """fp32 SIMT conv + pool kernels vs torch (channel-major input view and channel-major flatten output)."""
    from neural_sp_b200 import ops
    torch.manual_seed(1)
    dev = "cuda"
    B, T, Fq, CI, CO = 2, 21, 20, 3, 32
    x = torch.randn(B, T, CI * Fq, device=dev)             # reference view: [B,T,CI,F]
    w = torch.randn(CO, CI, 3, 3, device=dev) * 0.2
    b = torch.randn(CO, device=dev) * 0.1
    y = ops.conv3x3_relu(x, w, b, B, T, Fq, in_chmajor=True)
    ref = torch.relu(F.conv2d(x.view(B, T, CI, Fq).transpose(1, 2), w, b, padding=1))     # [B,CO,T,F]
    assert (y - ref.permute(0, 2, 3, 1)).abs().max().item() <= 1e-4
    p = ops.maxpool2d(y, 2, 2, out_chmajor=True)
    pref = F.max_pool2d(ref, 2, 2, 0, ceil_mode=True)
    pref = pref.transpose(1, 2).reshape(B, pref.shape[2], -1)                             # conv.py:189 flatten
    assert (p - pref).abs().max().item() <= 1e-4
    pt = ops.maxpool_time(torch.randn(2, 9, 16, device=dev), 2)
    assert pt.shape == (2, 5, 16)


@pytest.mark.parametrize("kind", ["max_pool", "mean_pool", "drop", "add", "concat", "conv1d"])
@pytest.mark.parametrize("T", [40, 45])
def test_subsamplers_match_reference_formulas(kind, T):
    """The six intermediate subsamplers (subsampling.py:13-246) vs their torch formulas, incl. odd lengths."""
    import math
    from neural_sp_b200.encoders import subsampling as ss
    torch.manual_seed(T)
    dev, B, D, f = "cuda", 3, 64, 2
    xs = torch.randn(B, T, D, device=dev)
    xlens = torch.IntTensor([T, T - 3, 7])
    if kind == "max_pool":
        m = ss.MaxPoolSubsampler(f)
        ref = F.max_pool1d(xs.transpose(1, 2), f, f, 0, ceil_mode=True).transpose(1, 2)
        rl = [(n + 1 - f) // f + 1 for n in xlens.tolist()]
    elif kind == "mean_pool":
        m = ss.MeanPoolSubsampler(f)
        ref = F.avg_pool1d(xs.transpose(1, 2), f, f, 0, ceil_mode=True).transpose(1, 2)
        rl = [(n - f) // f + 1 for n in xlens.tolist()]      # the reference's floor formula for AvgPool1d (conv.py:446-450)
    elif kind == "drop":
        m = ss.DropSubsampler(f)
        ref = xs[:, ::f]
        rl = [max(1, math.ceil(n / f)) for n in xlens.tolist()]
    elif kind == "add":
        m = ss.AddSubsampler(f)
        xp = torch.cat([xs, xs.new_zeros(B, 1, D)], 1) if T % 2 else xs
        ref = xp[:, ::2][:, :(T + 1) // 2] + xp[:, 1::2]
        rl = [max(1, math.ceil(n / f)) for n in xlens.tolist()]
    elif kind == "concat":
        m = ss.ConcatSubsampler(f, D).to(dev)
        m.precision = "fp32"
        To = T // f
        ref = torch.relu(F.linear(xs[:, :To * f].reshape(B, To, f * D).double(), m.proj.weight.double(), m.proj.bias.double())).float()
        rl = [max(1, n // f) for n in xlens.tolist()]
    else:
        m = ss.Conv1dSubsampler(f, D).to(dev)
        m.precision = "fp32"
        ref = torch.relu(F.conv1d(xs.transpose(1, 2).double(), m.conv1d.weight.double(), m.conv1d.bias.double(), stride=f,
                                  padding=1).transpose(1, 2)).float()     # fp64: cuDNN conv defaults to TF32
        rl = [(n + 2 - 2 - 1) // f + 1 for n in xlens.tolist()]
    with torch.no_grad():
        y, yl = m(xs, xlens)
    assert y.shape == ref.shape, (y.shape, ref.shape)
    assert (y - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    assert [int(v) for v in yl] == rl
