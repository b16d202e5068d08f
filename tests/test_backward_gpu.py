"""Backward (training) path parity.

Unit level: every hand-written backward kernel against torch autograd of a plain fp32 torch restatement of the same
op on the same seeded inputs (the checker runs on the GPU in fp32/fp64; tolerances state the operand precision).
End to end: parameter gradients of whole encoders (CNN front-end + Conformer / Transformer blocks + max-pool +
final LayerNorm) against gradients of the UNMODIFIED reference (tests/golden/encgrad_*.npz, torch autograd on CPU)
for loss = sum(ys * w), w = 0 on padded frames:  fp32 mode 2e-3 of max|g| per tensor (north_star: 1e-3 rel fp32 on
activations; gradients pass through 2x as many GEMMs); bf16 mode 2e-1 (a wiring check: these are tiny models, d = 32..64,
where single bf16 roundings of ReLU / LayerNorm inputs move whole gradients -- the fp32 mode is the parity proof)."""
import glob
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, load_golden
from enc_util import build_ours

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def ops_():
    from neural_sp_b200 import ops
    return ops


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec,tol", [("bf16", 2e-2), ("tf32", 3e-3), ("fp32", 1e-4)])
@pytest.mark.parametrize("shape", [(1000, 256, 512), (77, 40, 24), (4100, 2048, 512), (11, 64, 64), (300, 1000, 136)])
def test_linear_wgrad(shape, prec, tol):
    ops = ops_()
    M, N, K = shape
    torch.manual_seed(0)
    dy = torch.randn(M, N, device=DEV)
    x = torch.randn(M, K, device=DEV)
    if prec == "bf16":
        dy, x = dy.bfloat16().float(), x.bfloat16().float()
    ref = (dy.double().t() @ x.double()) * 0.5
    dw = torch.zeros(N, K, device=DEV)
    ops.linear_wgrad(dy, x, prec, dw, alpha=0.5, accumulate=False)
    assert rel_err(dw, ref) <= tol
    ops.linear_wgrad(dy, x, prec, dw, alpha=0.5, accumulate=True)          # accumulation doubles it
    assert rel_err(dw, 2 * ref) <= tol


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_linear_save_pre(prec):
    ops = ops_()
    torch.manual_seed(1)
    x = torch.randn(300, 64, device=DEV)
    w = torch.randn(256, 64, device=DEV) / 8
    b = torch.randn(256, device=DEV)
    xin = x.bfloat16() if prec == "bf16" else x
    h, z = ops.linear(xin, ops.prepare_weight(w, prec), b, prec=prec, act="swish",
                      out_dtype=torch.bfloat16 if prec == "bf16" else torch.float32, save_pre=True)
    zr = x @ w.t() + b
    tol = 3e-2 if prec == "bf16" else 1e-4
    assert rel_err(z.float(), zr) <= tol and rel_err(h.float(), F.silu(zr)) <= tol
    g, pre = ops.linear(xin, ops.prepare_weight(w, prec), b, prec=prec, glu=True,
                        out_dtype=torch.bfloat16 if prec == "bf16" else torch.float32, save_pre=True)
    assert rel_err(pre.float(), zr) <= tol and rel_err(g.float(), F.glu(zr, dim=-1)) <= tol


@pytest.mark.parametrize("D", [64, 256, 512, 36])
def test_layernorm_bwd(D):
    ops = ops_()
    torch.manual_seed(2)
    M = 333
    x = torch.randn(M, D, device=DEV, requires_grad=True)
    gamma = torch.randn(D, device=DEV, requires_grad=True)
    beta = torch.randn(D, device=DEV, requires_grad=True)
    dy, dres = torch.randn(M, D, device=DEV), torch.randn(M, D, device=DEV)
    y = F.layer_norm(x, (D,), gamma, beta, 1e-12)
    y.backward(dy)
    dg, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    dcol = torch.ones(D, device=DEV)
    dx, dxb = ops.layernorm_bwd(dy, x.detach(), gamma.detach(), 1e-12, dres=dres, dgamma=dg, dbeta=db, want_bf16=True,
                                dcol=dcol, dcol_alpha=0.5)
    assert rel_err(dcol, 1 + 0.5 * (x.grad + dres).sum(0)) <= 2e-5          # fused bias gradient of the consuming branch
    assert rel_err(dx, x.grad + dres) <= 2e-5
    assert rel_err(dxb.float(), x.grad + dres) <= 1e-2
    assert rel_err(dg, gamma.grad) <= 2e-5 and rel_err(db, beta.grad) <= 2e-5


@pytest.mark.parametrize("act", ["relu", "swish", "gelu", "gelu_accurate"])
def test_act_and_glu_bwd(act):
    ops = ops_()
    torch.manual_seed(3)
    z = torch.randn(100, 96, device=DEV, requires_grad=True)
    dh = torch.randn(100, 96, device=DEV)
    fn = {"relu": F.relu, "swish": F.silu, "gelu": F.gelu,
          "gelu_accurate": lambda t: 0.5 * t * (1 + torch.tanh(0.7978845608028654 * (t + 0.044715 * t ** 3)))}[act]
    fn(z).backward(dh)
    assert rel_err(ops.act_bwd(dh, z.detach(), act), z.grad) <= 1e-5
    assert rel_err(ops.act_bwd(dh.bfloat16(), z.detach().bfloat16(), act).float(), z.grad) <= 3e-2
    pre = torch.randn(100, 128, device=DEV, requires_grad=True)
    dg = torch.randn(100, 64, device=DEV)
    F.glu(pre, dim=-1).backward(dg)
    assert rel_err(ops.glu_bwd(dg, pre.detach()), pre.grad) <= 1e-5
    # fused bias-gradient variants (bf16): result and its column sums
    db = torch.zeros(96, device=DEV)
    dzb = ops.act_bwd_bias(dh.bfloat16(), z.detach().bfloat16(), act, db)
    assert rel_err(dzb.float(), z.grad) <= 3e-2 and rel_err(db, z.grad.sum(0)) <= 2e-2
    dbg = torch.zeros(128, device=DEV)
    dpb = ops.act_bwd_bias(dg.bfloat16(), pre.detach().bfloat16(), None, dbg, glu=True)
    assert rel_err(dpb.float(), pre.grad) <= 3e-2 and rel_err(dbg, pre.grad.sum(0)) <= 2e-2


def test_colsum_pool_relu():
    ops = ops_()
    torch.manual_seed(4)
    x = torch.randn(1234, 100, device=DEV)
    y = torch.ones(100, device=DEV)
    ops.colsum_acc(x, y, alpha=0.5)
    assert rel_err(y, 1 + 0.5 * x.sum(0)) <= 1e-5
    yb = torch.zeros(100, device=DEV)
    ops.colsum_acc(x.bfloat16(), yb)
    assert rel_err(yb, x.bfloat16().float().sum(0)) <= 1e-4
    for T, f in ((50, 2), (51, 2), (37, 3)):
        xt = torch.randn(3, T, 24, device=DEV, requires_grad=True)
        yt = F.max_pool1d(xt.transpose(1, 2), f, f, 0, ceil_mode=True).transpose(1, 2)
        dy = torch.randn_like(yt)
        yt.backward(dy)
        assert torch.equal(ops.maxpool_time_bwd(xt.detach(), dy, f), xt.grad)
    a = torch.randn(2, 9, 11, 8, device=DEV).relu()
    dx = torch.randn_like(a)
    assert torch.equal(ops.relu_mask(dx, a), dx * (a > 0))


@pytest.mark.parametrize("pool,chmajor", [((2, 2), False), ((2, 2), True), ((1, 2), False), ((1, 1), True), ((3, 2), False)])
def test_maxpool2d_relu_bwd(pool, chmajor):
    ops = ops_()
    torch.manual_seed(5)
    B, T, Fq, C = 2, 13, 11, 8
    pre = torch.randn(B, C, T, Fq, device=DEV, requires_grad=True)           # reference layout [B,C,T,F]
    a = pre.relu()
    p = F.max_pool2d(a, pool, pool, 0, ceil_mode=True)
    dy_ref = torch.randn_like(p)
    p.backward(dy_ref)
    a_cl = a.detach().permute(0, 2, 3, 1).contiguous()
    if chmajor:     # [B, T', C*F']  (index c*F' + f)
        dy = dy_ref.permute(0, 2, 1, 3).reshape(B, p.shape[2], -1).contiguous()
    else:           # [B, T', F', C]
        dy = dy_ref.permute(0, 2, 3, 1).contiguous()
    dz = ops.maxpool2d_relu_bwd(a_cl, dy, pool[0], pool[1], in_chmajor=chmajor)
    assert torch.equal(dz, pre.grad.permute(0, 2, 3, 1).contiguous())
    if not chmajor:        # bf16 (vectorised) variant on bf16-exact inputs
        pre_b = pre.detach().bfloat16().float().requires_grad_(True)
        p_b = F.max_pool2d(pre_b.relu(), pool, pool, 0, ceil_mode=True)
        dy_b = dy_ref.bfloat16().float()
        p_b.backward(dy_b)
        dz_b = ops.maxpool2d_relu_bwd(pre_b.detach().relu().permute(0, 2, 3, 1).contiguous().bfloat16(),
                                      dy_b.permute(0, 2, 3, 1).contiguous().bfloat16(), pool[0], pool[1])
        assert torch.equal(dz_b.float(), pre_b.grad.permute(0, 2, 3, 1).contiguous())


@pytest.mark.parametrize("CI,first", [(1, True), (32, False)])
def test_conv3x3_wgrad_and_dgrad(CI, first):
    ops = ops_()
    torch.manual_seed(6)
    B, T, Fq, CO = 2, 21, 19, 32
    x = torch.randn(B, CI, T, Fq, device=DEV, requires_grad=True)
    w = (torch.randn(CO, CI, 3, 3, device=DEV) * 0.2).requires_grad_(True)
    b = torch.randn(CO, device=DEV, requires_grad=True)
    y = F.conv2d(x, w, b, padding=1)
    dz_ref = torch.randn_like(y)
    y.backward(dz_ref)
    dz = dz_ref.permute(0, 2, 3, 1).contiguous()                               # [B,T,F,CO]
    a = x.detach().permute(0, 2, 1, 3).contiguous() if first else x.detach().permute(0, 2, 3, 1).contiguous()
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    ops.conv3x3_wgrad(a, dz, dw, db, B, T, Fq, in_chmajor=first)
    assert rel_err(dw, w.grad) <= 2e-5 and rel_err(db, b.grad) <= 2e-5
    if not first:          # tcgen05 implicit-GEMM weight gradient (bf16 operands)
        dwt, dbt = torch.zeros_like(w), torch.zeros_like(b)
        ab, zb = a.bfloat16(), dz.bfloat16()
        ops.conv3x3_wgrad(ab, zb, dwt, dbt, B, T, Fq)
        ref_w = torch.autograd.grad(F.conv2d(ab.float().permute(0, 3, 1, 2), w, b, padding=1), w, zb.float().permute(0, 3, 1, 2))[0]
        assert rel_err(dwt, ref_w) <= 1e-2 and rel_err(dbt, zb.float().sum((0, 1, 2))) <= 1e-3
    if not first:
        wd = ops.conv3x3_dgrad_weight(w)
        dx = ops.conv3x3_relu(dz, wd, torch.zeros(CI, device=DEV), B, T, Fq, relu=False)
        assert rel_err(dx, x.grad.permute(0, 2, 3, 1)) <= 2e-5


# ---------------------------------------------------------------------------------------------
def _attn_ref(q, k, v, r, u, vb, klens, H, clamp, causal, lookahead):
    """fp32 torch restatement of the attention core (relative_multihead_attention.py:173-215)."""
    B, T, D = q.shape
    dk = D // H
    qh, kh, vh = (t.view(B, T, H, dk) for t in (q, k, v))
    ac = torch.einsum("bihd,bjhd->bhij", qh + (u if u is not None else 0), kh)
    e = ac
    if r is not None:
        i = torch.arange(T, device=q.device)
        dist = (i[:, None] - i[None, :]).abs()
        if clamp > 0:
            dist = dist.clamp(max=clamp)
        dist = dist.clamp(max=r.shape[0] - 1)
        rh = r.view(-1, H, dk)
        bd_raw = torch.einsum("bihd,nhd->bhin", qh + (vb if vb is not None else 0), rh)
        e = e + torch.gather(bd_raw, 3, dist[None, None].expand(B, H, T, T))
    e = e / dk ** 0.5
    j = torch.arange(T, device=q.device)
    mask = j[None, None, :] < klens[:, None, None].long()
    if causal:
        mask = mask & (j[None, None, :] <= (j[None, :, None] + lookahead))
    else:
        mask = mask.expand(B, T, T)
    e = e.masked_fill(~mask[:, None], torch.finfo(torch.float32).min)
    aw = e.softmax(-1)
    return torch.einsum("bhij,bjhd->bihd", aw, vh).reshape(B, T, D)


@pytest.mark.parametrize("cfg", [
    dict(T=70, H=4, dk=16, rel=True, clamp=10, xl=False, causal=False, bf16=False),
    dict(T=70, H=2, dk=64, rel=True, clamp=-1, xl=True, causal=False, bf16=False),
    dict(T=45, H=2, dk=64, rel=False, clamp=-1, xl=False, causal=True, bf16=False),
    dict(T=150, H=4, dk=64, rel=True, clamp=10, xl=False, causal=False, bf16=False),
    dict(T=150, H=4, dk=64, rel=True, clamp=10, xl=False, causal=False, bf16=True),
    dict(T=33, H=1, dk=128, rel=True, clamp=5, xl=True, causal=True, bf16=False),
    # tcgen05 backward envelope (bf16, d_k = 64, clamped / no relative term): one, two and three key tiles, masks
    dict(T=300, H=2, dk=64, rel=True, clamp=10, xl=False, causal=False, bf16=True),
    dict(T=128, H=1, dk=64, rel=False, clamp=-1, xl=False, causal=False, bf16=True),
    dict(T=260, H=2, dk=64, rel=True, clamp=3, xl=False, causal=True, bf16=True),
    dict(T=500, H=8, dk=64, rel=True, clamp=10, xl=False, causal=False, bf16=True),
])
def test_attention_bwd(cfg):
    ops = ops_()
    torch.manual_seed(7)
    B, T, H, dk = 3, cfg["T"], cfg["H"], cfg["dk"]
    D = H * dk
    klens = torch.tensor([T, max(1, T - 13), max(1, T // 2)], dtype=torch.int32, device=DEV)
    qkv = (torch.randn(B, T, 3 * D, device=DEV) * 0.7)
    nrows = (min(T, cfg["clamp"] + 1) if cfg["clamp"] > 0 else T) if cfg["rel"] else 0
    r = torch.randn(nrows, D, device=DEV) * 0.7 if cfg["rel"] else None
    u = torch.randn(H, dk, device=DEV) * 0.3 if cfg["xl"] else None
    vb = torch.randn(H, dk, device=DEV) * 0.3 if cfg["xl"] else None
    dout = torch.randn(B, T, D, device=DEV)
    dt = torch.bfloat16 if cfg["bf16"] else torch.float32
    if cfg["bf16"]:
        qkv, dout = qkv.bfloat16().float(), dout.bfloat16().float()
        r = r.bfloat16().float() if r is not None else None
    leaves = [t.clone().requires_grad_(True) if t is not None else None for t in (qkv, r, u, vb)]
    lq, lr, lu, lv = leaves
    out_ref = _attn_ref(lq[:, :, :D], lq[:, :, D:2 * D], lq[:, :, 2 * D:], lr, lu, lv, klens, H, cfg["clamp"], cfg["causal"], 1)
    out_ref.backward(dout)
    kw = dict(causal=cfg["causal"], lookahead=1 if cfg["causal"] else 0)
    qd = qkv.to(dt)
    rd = r.to(dt) if r is not None else None
    out, stats = ops.relpos_attention(qd[:, :, :D], qd[:, :, D:2 * D], qd[:, :, 2 * D:], klens, H, r=rd, u_bias=u, v_bias=vb,
                                      clamp_len=cfg["clamp"], want_stats=True, **kw)
    tc_case = cfg["bf16"] and dk == 64 and not cfg["xl"] and (not cfg["rel"] or 1 <= cfg["clamp"] <= 15)
    assert (stats is not None) == tc_case          # the tensor-core kernels ran exactly inside their envelope
    tol = 4e-2 if cfg["bf16"] else 2e-4
    assert rel_err(out.float(), out_ref.detach()) <= tol
    dr = torch.zeros(nrows, D, device=DEV) if cfg["rel"] else None
    du = torch.zeros(D, device=DEV) if cfg["xl"] else None
    dvb = torch.zeros(D, device=DEV) if cfg["xl"] else None
    dqkv = ops.relpos_attention_bwd(qd[:, :, :D], qd[:, :, D:2 * D], qd[:, :, 2 * D:], klens, H, out, dout.to(dt), r=rd,
                                    u_bias=u, v_bias=vb, clamp_len=cfg["clamp"], dr=dr, du=du, dvb=dvb, stats=stats, **kw)
    assert rel_err(dqkv.float(), lq.grad) <= tol
    if cfg["rel"]:
        assert rel_err(dr, lr.grad) <= tol
    if cfg["xl"]:
        assert rel_err(du.view(H, dk), lu.grad) <= tol and rel_err(dvb.view(H, dk), lv.grad) <= tol


@pytest.mark.parametrize("d,k,causal,bf16", [(64, 7, False, False), (256, 15, False, False), (512, 15, False, True),
                                             (144, 3, True, False), (512, 31, False, False)])
def test_conformer_conv_bwd(d, k, causal, bf16):
    ops = ops_()
    torch.manual_seed(8)
    B, T = 3, 77
    x = torch.randn(B, T, d, device=DEV)
    dy = torch.randn(B, T, d, device=DEV)
    if bf16:
        x, dy = x.bfloat16().float(), dy.bfloat16().float()
    w = (torch.randn(d, 1, k, device=DEV) * 0.3).requires_grad_(True)
    b = torch.randn(d, device=DEV, requires_grad=True)
    g = (1 + 0.1 * torch.randn(d, device=DEV)).requires_grad_(True)
    be = (0.1 * torch.randn(d, device=DEV)).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    pad = k - 1 if causal else (k - 1) // 2
    z = F.conv1d(xr.transpose(1, 2), w, b, padding=pad, groups=d)
    if causal:
        z = z[:, :, :-pad]
    y = F.silu(F.layer_norm(z.transpose(1, 2), (d,), g, be, 1e-12))
    y.backward(dy)
    taps = w.detach().reshape(d, k).t().contiguous()
    dt = torch.bfloat16 if bf16 else torch.float32
    yk = ops.conformer_conv(x.to(dt), taps, b.detach(), "layer_norm", g.detach(), be.detach(), 1e-12, causal=causal)
    tol = 4e-2 if bf16 else 1e-4
    assert rel_err(yk.float(), y.detach()) <= tol
    dtaps, db, dg, dbe = torch.zeros_like(taps), torch.zeros(d, device=DEV), torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    dx = ops.conformer_conv_bwd(x.to(dt), taps, b.detach(), g.detach(), be.detach(), 1e-12, dy.to(dt), dtaps, db, dg, dbe,
                                causal=causal)
    assert rel_err(dx.float(), xr.grad) <= tol
    assert rel_err(dtaps.t().reshape(d, 1, k), w.grad) <= tol
    assert rel_err(db, b.grad) <= tol and rel_err(dg, g.grad) <= tol and rel_err(dbe, be.grad) <= tol


# ---------------------------------------------------------------------------------------------
GRAD_CASES = sorted(os.path.basename(f)[len("encgrad_"):-4] for f in glob.glob(os.path.join(GOLDEN, "encgrad_*.npz")))


def _loss_weights(shape, xlens_out, seed=4321):
    """Same projection as tests/golden/gen_golden_encoder.py::grad_loss_weights (zero on padded output frames)."""
    w = np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
    for b, n in enumerate(xlens_out):
        w[b, int(n):] = 0.0
    return w


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-3), ("bf16", 2e-1)])
@pytest.mark.parametrize("name", GRAD_CASES)
def test_encoder_param_grads_match_reference(name, precision, tol):
    g = load_golden("enc_%s.npz" % name)
    gg = load_golden("encgrad_%s.npz" % name)
    dev = torch.device(DEV)
    enc = build_ours(g, dev, precision)
    enc.train()
    out = enc(torch.from_numpy(g["xs"]).to(dev), torch.IntTensor(g["xlens"].tolist()), task="all")
    ys = out["ys"]["xs"]
    assert ys.requires_grad
    w = torch.from_numpy(_loss_weights(tuple(ys.shape), out["ys"]["xlens"].tolist())).to(dev)
    loss = (ys * w).sum()
    ftol = 1e-3 if precision == "fp32" else 5e-2
    assert abs(float(loss.detach()) - float(gg["loss"])) <= ftol * max(1.0, float(w.abs().sum()) * 0.01)
    loss.backward()
    # error of each tensor relative to its own max |g|, floored at 1e-3 of the largest gradient of the model (some
    # gradients are analytically zero, e.g. key biases under the softmax's shift invariance)
    gmax = max(float(np.abs(gg[k]).max()) for k in gg.files if k.startswith("g."))
    bad = []
    for k, p in enc.named_parameters():
        ref = torch.from_numpy(gg["g." + k]).double()
        assert p.grad is not None, k
        e = float((p.grad.detach().cpu().double() - ref).abs().max() / max(float(ref.abs().max()), 1e-3 * gmax))
        if not e <= tol:
            bad.append((k, e))
    assert not bad, (name, precision, bad[:10], len(bad))
    # forward in train mode equals the inference path (eval) on the same weights
    enc.eval()
    with torch.no_grad():
        ys_eval = enc(torch.from_numpy(g["xs"]).to(dev), torch.IntTensor(g["xlens"].tolist()), task="all")["ys"]["xs"]
    assert rel_err(ys.detach(), ys_eval) <= (1e-5 if precision == "fp32" else 3e-2)


def test_ctc_training_step_end_to_end():
    """Encoder + CTC head + CTC loss: loss.backward() fills every parameter's .grad through the CUDA path."""
    from neural_sp_b200.decoders.ctc import CTC
    g = load_golden("enc_conformer_small.npz")
    dev = torch.device(DEV)
    enc = build_ours(g, dev, "bf16")
    enc.train()
    ctc = CTC(eos=2, blank=0, enc_n_units=64, vocab=40, lsm_prob=0.1, fc_list="32").to(dev)
    ctc.train()
    out = enc(torch.from_numpy(g["xs"]).to(dev), torch.IntTensor(g["xlens"].tolist()), task="ys")
    ys = [[5, 6, 7, 8], [9, 10, 11], [12, 13]]
    loss, _ = ctc(out["ys"]["xs"], out["ys"]["xlens"], ys)
    loss.backward()
    for k, p in list(enc.named_parameters()) + list(ctc.named_parameters()):
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        assert float(p.grad.abs().max()) > 0, k


# ---- streaming conv-module kernels (csrc/conformer_conv_stream.cu): every covered (d, k), both dtypes, causal, ragged T ----
@pytest.mark.parametrize("d,k,T,causal,bf16", [(512, 15, 125, False, True), (512, 15, 500, False, True), (512, 15, 77, False, False),
                                               (256, 15, 250, False, True), (256, 7, 33, True, False), (128, 5, 64, False, True),
                                               (64, 3, 50, True, False), (512, 7, 31, True, True), (128, 15, 200, False, False)])
def test_conv_module_streaming_kernels(d, k, T, causal, bf16):
    """Forward, dx and the four parameter gradients of Swish(LayerNorm(dwconv(x))) against torch autograd in fp64."""
    ops = ops_()
    torch.manual_seed(d + k + T)
    B = 3
    x = torch.randn(B, T, d, device=DEV)
    dy = torch.randn(B, T, d, device=DEV)
    if bf16:
        x, dy = x.bfloat16().float(), dy.bfloat16().float()
    w = (torch.randn(d, 1, k, device=DEV) * 0.3).double().requires_grad_(True)
    b = torch.randn(d, device=DEV).double().requires_grad_(True)
    g = (1 + 0.1 * torch.randn(d, device=DEV)).double().requires_grad_(True)
    be = (0.1 * torch.randn(d, device=DEV)).double().requires_grad_(True)
    xr = x.double().requires_grad_(True)
    pad = k - 1 if causal else (k - 1) // 2
    z = F.conv1d(xr.transpose(1, 2), w, b, padding=pad, groups=d)
    if causal:
        z = z[:, :, :-pad]
    y = F.silu(F.layer_norm(z.transpose(1, 2), (d,), g, be, 1e-12))
    y.backward(dy.double())
    taps = w.detach().float().reshape(d, k).t().contiguous()
    dt = torch.bfloat16 if bf16 else torch.float32
    f32 = lambda t: t.detach().float()
    yk = ops.conformer_conv(x.to(dt), taps, f32(b), "layer_norm", f32(g), f32(be), 1e-12, causal=causal)
    tol = 3e-2 if bf16 else 1e-4
    assert rel_err(yk.float(), f32(y)) <= tol
    dtaps, db, dg, dbe = torch.zeros_like(taps), torch.zeros(d, device=DEV), torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    dx = ops.conformer_conv_bwd(x.to(dt), taps, f32(b), f32(g), f32(be), 1e-12, dy.to(dt), dtaps, db, dg, dbe, causal=causal)
    assert rel_err(dx.float(), f32(xr.grad)) <= tol
    assert rel_err(dtaps.t().reshape(d, 1, k), f32(w.grad)) <= tol
    assert rel_err(db, f32(b.grad)) <= tol and rel_err(dg, f32(g.grad)) <= tol and rel_err(dbe, f32(be.grad)) <= tol
    # accumulate semantics (+=) of the parameter gradients: a second call doubles them
    ops.conformer_conv_bwd(x.to(dt), taps, f32(b), f32(g), f32(be), 1e-12, dy.to(dt), dtaps, db, dg, dbe, causal=causal)
    assert rel_err(0.5 * dg, f32(g.grad)) <= tol and rel_err(0.5 * dtaps.t().reshape(d, 1, k), f32(w.grad)) <= tol
