"""TMA-store epilogue variant of the tcgen05 GEMM (csrc/gemm_tma_epi.cu, nsp_set_gemm_epilogue(1)) against the same
PyTorch references as tests/test_gemm_gpu.py.  The variant is opt-in until these tests have passed on a B200
(NSP_EXPERIMENTAL=1 python -m pytest tests/test_gemm_tma_epilogue_gpu.py -m gpu); afterwards the whole GPU suite can be
re-run with NSP_GEMM_EPILOGUE=tma to exercise it under every caller."""
import pytest
import torch

from test_gemm_gpu import _ref

pytestmark = pytest.mark.gpu


def _mode(mode):
    from neural_sp_b200 import _lib
    old = _lib.lib.nsp_get_gemm_epilogue()
    _lib.check(_lib.lib.nsp_set_gemm_epilogue(mode), "nsp_set_gemm_epilogue")
    yield _lib.lib
    _lib.lib.nsp_set_gemm_epilogue(old)


@pytest.fixture
def tma_epilogue():
    yield from _mode(1)


@pytest.fixture
def cta_pairs():
    yield from _mode(2)


def _run(lib, M, N, K, bias=True, act=None, glu=False, residual=False, alpha=1.0, out_dtype=torch.float32, save_pre=False,
         expect_tma=True, expect_pair=False):
    from neural_sp_b200 import ops
    torch.manual_seed(M * 7 + N * 3 + K)
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda") if bias else None
    nout = N // 2 if glu else N
    r = torch.randn(M, nout, device="cuda") if residual else None
    xr, wr = x.bfloat16().float(), w.bfloat16().float()
    ref = _ref(xr, wr, b, act, glu, r, alpha)
    before, before2 = lib.nsp_gemm_tma_epilogue_launches(), lib.nsp_gemm_cta_pair_launches()
    res = ops.linear(x, ops.prepare_weight(w, "bf16"), b, prec="bf16", act=act, glu=glu, residual=r, alpha=alpha,
                     out_dtype=out_dtype, save_pre=save_pre)
    torch.cuda.synchronize()
    assert (lib.nsp_gemm_tma_epilogue_launches() - before == 1) == expect_tma, "envelope routing"
    assert (lib.nsp_gemm_cta_pair_launches() - before2 == 1) == expect_pair, "CTA-pair routing"
    out, pre = res if save_pre else (res, None)
    assert out.shape == (M, nout)
    tol = 8e-3 if out_dtype == torch.bfloat16 else 1e-4
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= tol, (M, N, K, act, glu, residual, err)
    if save_pre:
        z = xr.double() @ wr.double().t() + (b.double() if b is not None else 0.0)
        assert pre.shape == (M, N) and pre.dtype == torch.bfloat16
        assert (pre.double() - z).abs().max().item() <= 8e-3 * z.abs().max().item()
    # untouched neighbours: a second, smaller problem in a larger buffer must not be written outside [M, nout]
    return out


@pytest.mark.parametrize("M", [16000, 16037])
def test_plain_fp32_and_bf16_outputs(tma_epilogue, M):
    _run(tma_epilogue, M, 512, 512)                                        # BN = 128, fp32 out, two chunks per warp
    _run(tma_epilogue, M, 512, 512, out_dtype=torch.bfloat16, bias=False)  # SWIZZLE_64B staging
    _run(tma_epilogue, M, 2048, 512, out_dtype=torch.bfloat16)             # BN = 256, four chunks per warp
    _run(tma_epilogue, M, 2048, 256)                                       # BN = 256, fp32 out (3-stage ring)


@pytest.mark.parametrize("M", [16000, 16037])
def test_activations_and_saved_preactivation(tma_epilogue, M):
    _run(tma_epilogue, M, 2048, 512, act="swish", out_dtype=torch.bfloat16, save_pre=True)     # FFN-in of the training path
    _run(tma_epilogue, M, 512, 256, act="relu")
    _run(tma_epilogue, M, 1024, 512, glu=True, out_dtype=torch.bfloat16, save_pre=True)        # conv module pointwise-1
    _run(tma_epilogue, M, 1024, 512, glu=True)


@pytest.mark.parametrize("M", [16000, 16037])
def test_residual_epilogue(tma_epilogue, M):
    _run(tma_epilogue, M, 512, 2048, residual=True, alpha=0.5)            # FFN-out: residual tile fetched under the mainloop
    _run(tma_epilogue, M, 512, 512, residual=True, bias=False)
    _run(tma_epilogue, 20000, 576, 512, residual=True)                    # last column tile half empty (576 = 4.5 x 128)


def test_inplace_residual(tma_epilogue):
    """out aliases the residual (the inference path updates the residual stream in place)."""
    from neural_sp_b200 import ops
    torch.manual_seed(3)
    x = torch.randn(16037, 512, device="cuda")
    w = torch.randn(512, 512, device="cuda") / 512 ** 0.5
    res = torch.randn(16037, 512, device="cuda")
    ref = res.double() + 0.5 * (x.bfloat16().double() @ w.bfloat16().double().t())
    before = tma_epilogue.nsp_gemm_tma_epilogue_launches()
    out = ops.linear(x, ops.prepare_weight(w, "bf16"), None, prec="bf16", residual=res, alpha=0.5, out=res)
    torch.cuda.synchronize()
    assert tma_epilogue.nsp_gemm_tma_epilogue_launches() - before == 1
    assert out.data_ptr() == res.data_ptr()
    assert (out.double() - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()


def test_outside_the_envelope_falls_back(tma_epilogue):
    _run(tma_epilogue, 500, 1024, 256, act="swish", expect_tma=False)      # few tiles -> 64-wide direct-store kernel
    _run(tma_epilogue, 16000, 1000, 512, expect_tma=False)                 # output width not a multiple of 32
    _run(tma_epilogue, 129, 136, 72, expect_tma=False)


def test_strided_output_views(tma_epilogue):
    """Output written into a column slice of a wider buffer (pitch != width): neighbours stay untouched."""
    from neural_sp_b200 import _lib, ops
    torch.manual_seed(5)
    M, N, K = 16037, 512, 512
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    buf = torch.full((M, 3 * N), 7.0, device="cuda")
    out = buf[:, N:2 * N]
    ops.linear(x, ops.prepare_weight(w, "bf16"), None, prec="bf16", out=out)
    torch.cuda.synchronize()
    ref = x.bfloat16().double() @ w.bfloat16().double().t()
    assert (out.double() - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    assert torch.all(buf[:, :N] == 7.0) and torch.all(buf[:, 2 * N:] == 7.0)


# ---------------------------------------------------------------------------------------------
# mode 2: CTA pairs (cta_group::2), 256-row tiles.  Run these only after the mode-1 tests above are green.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M", [16000, 16037, 16200])             # 16200: the second CTA of the last pair has no live rows
def test_cta_pairs_plain(cta_pairs, M):
    _run(cta_pairs, M, 512, 512, expect_pair=True)                                         # 256 x 256 tiles, fp32 out
    _run(cta_pairs, M, 2048, 512, out_dtype=torch.bfloat16, bias=False, expect_pair=True)
    _run(cta_pairs, M, 1536, 512, out_dtype=torch.bfloat16, expect_pair=True)              # fused QKV projection
    _run(cta_pairs, M, 640, 512, expect_pair=True)                                         # 256 x 128 tiles (640 % 256 != 0)


def test_cta_pairs_epilogues(cta_pairs):
    _run(cta_pairs, 16037, 2048, 512, act="swish", out_dtype=torch.bfloat16, save_pre=True, expect_pair=True)
    _run(cta_pairs, 16037, 512, 2048, residual=True, alpha=0.5, expect_pair=True)          # residual: 256 x 128 tiles
    _run(cta_pairs, 8000, 512, 512, residual=True, expect_pair=True)
    _run(cta_pairs, 8000, 2048, 512, act="relu", expect_pair=True)
    _run(cta_pairs, 16037, 1024, 512, glu=True, out_dtype=torch.bfloat16, save_pre=True)   # GLU stays single-CTA
    _run(cta_pairs, 4000, 512, 2048, residual=True, expect_pair=True)                      # 64 pair tiles: one 86 % wave
    _run(cta_pairs, 2000, 512, 2048, residual=True, expect_tma=False)                      # too few tiles for pairs and for 128-wide tiles


def test_cta_pairs_long_k_many_tiles_per_pair(cta_pairs):
    """Several tiles per pair and K = 4096: exercises ring wrap-around, both accumulator stages and the phase bits."""
    _run(cta_pairs, 40000, 1024, 4096, expect_pair=True)


# ---------------------------------------------------------------------------------------------
# weight gradient: staged cp.reduce.async.bulk.tensor epilogue (modes >= 1)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(4000, 512, 2048), (4000, 2048, 512), (16000, 512, 512), (1000, 256, 512), (77, 40, 24),
                                   (4100, 2048, 512), (11, 64, 64), (300, 1000, 136), (16037, 1536, 512)])
def test_wgrad_bulk_reduce_epilogue(tma_epilogue, shape):
    from neural_sp_b200 import ops
    M, N, K = shape
    torch.manual_seed(0)
    dy = torch.randn(M, N, device="cuda").bfloat16().float()
    x = torch.randn(M, K, device="cuda").bfloat16().float()
    ref = (dy.double().t() @ x.double()) * 0.5
    dw = torch.zeros(N, K, device="cuda")
    before = tma_epilogue.nsp_wgrad_tma_epilogue_launches()
    ops.linear_wgrad(dy, x, "bf16", dw, alpha=0.5, accumulate=False)
    torch.cuda.synchronize()
    assert tma_epilogue.nsp_wgrad_tma_epilogue_launches() - before == 1
    err = float((dw.double() - ref).abs().max() / ref.abs().max())
    assert err <= 2e-2, (shape, err)
    ops.linear_wgrad(dy, x, "bf16", dw, alpha=0.5, accumulate=True)          # accumulation doubles it
    torch.cuda.synchronize()
    err = float((dw.double() - 2 * ref).abs().max() / (2 * ref).abs().max())
    assert err <= 2e-2, (shape, err)


def test_wgrad_bulk_reduce_into_a_row_block_of_a_fused_gradient(tma_epilogue):
    """dW written into rows [512, 1024) of a [1536, 512] fused-QKV gradient view: neighbours untouched."""
    from neural_sp_b200 import ops
    torch.manual_seed(1)
    M, N, K = 4000, 512, 512
    dy = torch.randn(M, N, device="cuda").bfloat16().float()
    x = torch.randn(M, K, device="cuda").bfloat16().float()
    full = torch.full((3 * N, K), 3.0, device="cuda")
    ops.linear_wgrad(dy, x, "bf16", full[N:2 * N], accumulate=True)
    torch.cuda.synchronize()
    ref = 3.0 + dy.double().t() @ x.double()
    assert float((full[N:2 * N].double() - ref).abs().max() / ref.abs().max()) <= 2e-2
    assert torch.all(full[:N] == 3.0) and torch.all(full[2 * N:] == 3.0)


# ---- the round-1 direct-store family stays selectable (NSP_GEMM_EPILOGUE=direct) and correct: since round 2 the library
# ---- default is the TMA-store / CTA-pair family, so these shapes pin the older kernels explicitly
@pytest.fixture
def direct_store():
    yield from _mode(0)


@pytest.mark.parametrize("M", [16000, 4001])
def test_direct_store_family_still_correct(direct_store, M):
    _run(direct_store, M, 512, 512, expect_tma=False)
    _run(direct_store, M, 2048, 512, act="swish", out_dtype=torch.bfloat16, save_pre=True, expect_tma=False)
    _run(direct_store, M, 512, 2048, residual=True, alpha=0.5, expect_tma=False)
    _run(direct_store, M, 1024, 512, glu=True, out_dtype=torch.bfloat16, expect_tma=False)


def test_default_family_is_tma_pairs():
    from neural_sp_b200 import _lib
    import os
    if not os.environ.get("NSP_GEMM_EPILOGUE"):
        assert _lib.lib.nsp_get_gemm_epilogue() == 2
