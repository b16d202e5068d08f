// Weight-gradient GEMM on tcgen05 (sm_100a):   dW[N,K] (+)= alpha * dY[M,N]^T * X[M,K]
//
// The reference gets this from autograd of every nn.Linear / 1x1 nn.Conv1d on the encoder path
// (modules/positionwise_feed_forward.py:47-48, modules/relative_multihead_attention.py:57-60,
//  modules/conformer_convolution.py:44-69, encoders/conv.py:87, decoders/ctc.py:81-91).
//
// Both operands are read exactly as they sit in HBM (row-major [M, *], the reduction index M outermost), i.e. as
// MN-major tcgen05 operands: a TMA box of 64 rows x 128 bytes lands as one 8 KiB block of eight 8-row swizzle atoms,
// the shared-memory descriptors walk 128-byte column chunks with LBO = 8 KiB and 8-row groups with SBO = 1 KiB, and the
// instruction descriptor has both "major" bits set -- no transposed copies of activations are ever made.
// The reduction dimension (M = B*T' frames) is long while dW is small, so the tiles are split along M across CTAs
// (split-K) and reduced with vector red.global.add into the fp32 gradient buffer, which is also how gradient
// accumulation across calls comes for free.
//   warp 0: TMA producer | warp 1: single-thread MMA issuer | warps 2-5: epilogue (tcgen05.ld -> red.global.add.v4.f32)
// Epilogue mode 1 (nsp_set_gemm_epilogue >= 1, opt-in): the per-thread vector reds touch 32 lines per warp instruction
// (~8200 L1 wavefronts for a 128 x 256 tile, several times the mainloop of a short split), so the tile is instead staged as
// swizzled 32 x 32 fp32 sub-tiles in the (by then idle) operand ring and added by cp.reduce.async.bulk.tensor (L2 adds).
#include "tc_common.cuh"

namespace nsp {
namespace {

constexpr int WG_BM = 128;     // rows of dW per tile (= columns of dY)
constexpr int WG_ROWS = 64;    // reduction rows (frames) per pipeline stage
constexpr int WG_BOX = WG_ROWS * 128;   // bytes of one TMA box (64 rows x 128 B)
constexpr int WG_MAX_SEG = 3;

struct WgradMaps {
    CUtensorMap a[WG_MAX_SEG];   // dY (hi / lo)
    CUtensorMap b[WG_MAX_SEG];   // X  (hi / lo)
    CUtensorMap dw;              // epilogue mode 1: fp32 [N, K], box 32 x 32, SWIZZLE_128B
};

struct WgradArgs {
    int M, N, K, nseg, splits;
    float alpha;
    float* dw;
    int64_t lddw;
};

// MN-major operand, 128-byte swizzle: LBO = stride between 128-byte column chunks, SBO = stride between 8-row groups
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(WG_BOX >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <typename TIn, int BN, int STAGES, bool TMAEPI>
__global__ void __launch_bounds__(192, 1) wgrad_kernel(const __grid_constant__ WgradMaps maps, const WgradArgs g) {
    pdl_launch_dependents();      // PDL: the next kernel may start its prologue; ours overlaps the previous kernel's tail
    constexpr bool kBF16 = sizeof(TIn) == 2;
    constexpr int CHUNK = 128 / (int)sizeof(TIn);          // elements per 128-byte column chunk: 64 / 32
    constexpr int A_BOXES = WG_BM / CHUNK, B_BOXES = BN / CHUNK;
    constexpr int A_BYTES = A_BOXES * WG_BOX, B_BYTES = B_BOXES * WG_BOX;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int UK = 32 / (int)sizeof(TIn);              // reduction rows per MMA: 16 / 8
    constexpr int KSTEP_BYTES = (UK / 8) * 1024;           // 8-row groups per MMA x 1 KiB
    constexpr int TMEM_COLS = BN;
    static_assert(BN == 128 || BN == 256, "tile width");

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tfull_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (g.N + WG_BM - 1) / WG_BM;
    const int k_tiles = (g.K + BN - 1) / BN;
    const int tile = blockIdx.x % (n_tiles * k_tiles);
    const int split = blockIdx.x / (n_tiles * k_tiles);
    const int n_blk = tile % n_tiles, k_blk = tile / n_tiles;
    const int chunks = (g.M + WG_ROWS - 1) / WG_ROWS;
    const int per = (chunks + g.splits - 1) / g.splits;
    const int c_begin = split * per;
    const int c_end = min(chunks, c_begin + per);
    if (c_begin >= c_end) return;                          // uniform for the whole CTA

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < g.nseg; ++s) { tc::tma_prefetch_desc(&maps.a[s]); tc::tma_prefetch_desc(&maps.b[s]); }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
        tc::mbar_init(tfull_bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<TMEM_COLS>(tmem_holder);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();                    // the previous grid is complete: operands / residuals / outputs may be touched from here

    if (warp == 0) {
        if (tc::elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int s = 0; s < g.nseg; ++s) {
                for (int c = c_begin; c < c_end; ++c) {
                    tc::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    tc::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
#pragma unroll
                    for (int i = 0; i < A_BOXES; ++i)
                        tc::tma_load_2d(sa + i * WG_BOX, &maps.a[s], &full_bar[stage], n_blk * WG_BM + i * CHUNK, c * WG_ROWS);
#pragma unroll
                    for (int i = 0; i < B_BOXES; ++i)
                        tc::tma_load_2d(sb + i * WG_BOX, &maps.b[s], &full_bar[stage], k_blk * BN + i * CHUNK, c * WG_ROWS);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            // both operands MN-major: bits 15 (A) and 16 (B) of the instruction descriptor
            constexpr uint32_t idesc = tc::make_idesc(kBF16 ? 1u : 2u, WG_BM, BN) | (1u << 15) | (1u << 16);
            int stage = 0; uint32_t phase = 0;
            uint32_t accum = 0;
            for (int s = 0; s < g.nseg; ++s) {
                for (int c = c_begin; c < c_end; ++c) {
                    tc::mbar_wait(&full_bar[stage], phase);
                    tc::tc_fence_after();
                    const uint32_t sa = tc::smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t adesc = make_desc_mn_sw128(sa);
                    const uint64_t bdesc = make_desc_mn_sw128(sa + A_BYTES);
#pragma unroll
                    for (int k = 0; k < WG_ROWS / UK; ++k) {
                        const uint64_t koff = (uint64_t)((k * KSTEP_BYTES) >> 4);
                        if constexpr (kBF16) tc::umma_f16(tmem_base, adesc + koff, bdesc + koff, idesc, accum);
                        else tc::umma_tf32(tmem_base, adesc + koff, bdesc + koff, idesc, accum);
                        accum = 1;
                    }
                    tc::umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            tc::umma_commit(tfull_bar);
        }
    } else {
        // ===================== epilogue warps 2..5 =====================
        const int q = warp & 3;
        tc::mbar_wait_epi(tfull_bar, 0);
        tc::tc_fence_after();
        const int row = n_blk * WG_BM + q * 32 + lane;           // row of dW
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        if constexpr (TMAEPI) {
            // every MMA has retired (tfull), so every stage of the operand ring has been consumed: two 4 KiB staging slots
            // per warp alias the start of the ring
            uint8_t* stg = smem + (warp - 2) * 8192;
            const int row0 = n_blk * WG_BM + q * 32;
            int slot = 0;
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t r[32];
                tc::tmem_ld_32x32(t_row + (uint32_t)c, r);
                tc::tmem_ld_wait();
                const int col0 = k_blk * BN + c;
                if (row0 < g.N && col0 < g.K) {                  // warp-uniform; the bulk reduce clips partial sub-tiles
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = g.alpha * __uint_as_float(r[j]);
                    if (tc::elect_sync()) tc::bulk_wait_read<1>();      // the reduce that read this slot two sub-tiles ago is done
                    __syncwarp();
                    tc::st_row_f32(stg + slot * 4096, lane, v);
                    tc::fence_proxy_async_smem();
                    __syncwarp();
                    if (tc::elect_one()) { tc::tma_reduce_add_2d(&maps.dw, stg + slot * 4096, col0, row0); tc::bulk_commit(); }
                    slot ^= 1;
                }
            }
            if (tc::elect_sync()) tc::bulk_wait_read<0>();
            tc::tc_fence_before();
        } else {
        const bool vec_ok = (g.lddw % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.dw) & 15) == 0);
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            uint32_t r[32];
            tc::tmem_ld_32x32(t_row + (uint32_t)c, r);
            tc::tmem_ld_wait();
            const int col0 = k_blk * BN + c;
            if (row < g.N && col0 < g.K) {
                float* o = g.dw + (int64_t)row * g.lddw + col0;
                if (vec_ok && col0 + 32 <= g.K) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        red_add_v4(o + j, g.alpha * __uint_as_float(r[j]), g.alpha * __uint_as_float(r[j + 1]),
                                   g.alpha * __uint_as_float(r[j + 2]), g.alpha * __uint_as_float(r[j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j < g.K) atomicAdd(o + j, g.alpha * __uint_as_float(r[j]));
                }
            }
        }
        tc::tc_fence_before();
        }
    }

    __syncthreads();
    if (warp == 2) {
        tc::tc_fence_after();
        tc::tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

long long g_wgrad_tma_launches = 0;

template <typename TIn, int BN, int STAGES, bool TMAEPI>
nsp_status launch_wgrad(const WgradMaps& maps, WgradArgs& g, cudaStream_t st) {
    constexpr int CHUNK = 128 / (int)sizeof(TIn);
    constexpr size_t stage_bytes = (size_t)(WG_BM / CHUNK + BN / CHUNK) * WG_BOX;
    constexpr size_t smem = STAGES * stage_bytes + 1024 + 256;
    static_assert(smem <= 227 * 1024, "wgrad: shared memory budget");
    static_assert(!TMAEPI || stage_bytes >= 4 * 8192, "staging slots alias the first stage");
    auto kern = wgrad_kernel<TIn, BN, STAGES, TMAEPI>;
    static bool attr_set = false;
    if (!attr_set) {
        NSP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int tiles = ceil_div(g.N, WG_BM) * ceil_div(g.K, BN);
    const int chunks = ceil_div(g.M, WG_ROWS);
    int splits = ceil_div(2 * num_sms(), tiles);
    const int max_splits = ceil_div(chunks, 4);            // at least 4 pipeline stages of work per CTA
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    g.splits = splits;
    launch_k(kern, dim3((unsigned)(tiles * splits)), dim3(192), smem, st, maps, g);
    NSP_LAUNCH_OK();
    if (TMAEPI) ++g_wgrad_tma_launches;
    return NSP_OK;
}

}  // namespace

int gemm_epilogue_mode();                                  // gemm_tcgen05.cu
long long wgrad_tma_launch_count() { return g_wgrad_tma_launches; }

}  // namespace nsp

using namespace nsp;

extern "C" nsp_status nsp_linear_wgrad(int prec, const void* dy, const void* dy_lo, int64_t lddy,
                                       const void* x, const void* x_lo, int64_t ldx, int M, int N, int K,
                                       float alpha, float* dw, int64_t lddw, int accumulate, void* stream) {
    NSP_CHECK_ARG(dy && x && dw, "linear_wgrad: null pointer");
    NSP_CHECK_ARG(M > 0 && N > 0 && K > 0, "linear_wgrad: bad shape M=%d N=%d K=%d", M, N, K);
    NSP_CHECK_ARG(prec >= 0 && prec <= 2, "linear_wgrad: precision=%d", prec);
    NSP_CHECK_ARG(prec != 2 || (dy_lo && x_lo), "linear_wgrad: fp32 mode needs the lo splits");
    cudaStream_t st = (cudaStream_t)stream;
    const bool bf16 = prec == 0;
    const int es = bf16 ? 2 : 4;
    if (!accumulate) {
        if (lddw == K) NSP_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)N * K, st));
        else NSP_CUDA_OK(cudaMemset2DAsync(dw, sizeof(float) * lddw, 0, sizeof(float) * K, N, st));
    }
    WgradMaps maps;
    WgradArgs g;
    g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.dw = dw; g.lddw = lddw; g.splits = 1;
    g.nseg = prec == 2 ? 3 : 1;
    const void* as[3] = {dy, dy_lo, dy};
    const void* bs[3] = {x, x, x_lo};
    for (int s = 0; s < g.nseg; ++s) {
        if (!make_tmap_2d(&maps.a[s], as[s], es, bf16, (uint64_t)M, (uint64_t)N, (uint64_t)lddy, WG_ROWS)) return NSP_ERR_INVALID;
        if (!make_tmap_2d(&maps.b[s], bs[s], es, bf16, (uint64_t)M, (uint64_t)K, (uint64_t)ldx, WG_ROWS)) return NSP_ERR_INVALID;
    }
    if (bf16) {
        // opt-in bulk-reduce epilogue: needs a 16-byte aligned gradient buffer with a 16-byte multiple pitch
        const bool tma_epi = gemm_epilogue_mode() >= 1 && ((uintptr_t)dw % 16 == 0) && (lddw % 4 == 0);
        if (tma_epi) {
            if (!encode_tmap_2d(&maps.dw, dw, false, (uint64_t)N, (uint64_t)K, (uint64_t)lddw, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B, "dw"))
                return NSP_ERR_INVALID;
            if (K > 128) return launch_wgrad<__nv_bfloat16, 256, 4, true>(maps, g, st);
            return launch_wgrad<__nv_bfloat16, 128, 6, true>(maps, g, st);
        }
        if (K > 128) return launch_wgrad<__nv_bfloat16, 256, 4, false>(maps, g, st);
        return launch_wgrad<__nv_bfloat16, 128, 6, false>(maps, g, st);
    }
    // tf32 operands: the MN-major path needs the 128B_BASE32B swizzle atom (not wired up); the host side computes
    // parity-mode weight gradients on the K-major GEMM instead.
    set_error("linear_wgrad: only NSP_PREC_BF16 runs on the MN-major kernel; use nsp_linear_fwd on transposed operands for tf32/fp32");
    return NSP_ERR_UNSUPPORTED;
}
