// tcgen05 GEMM with a shared-memory staged, TMA bulk-store epilogue (bf16 operands):
//     out = residual + alpha * act(A[M,K] * W[N,K]^T + bias)        (act: none / relu / swish / gelu / GLU)
//
// Same contract, mainloop and warp roles as gemm_tcgen05.cu (which it replaces for the large bf16 GEMMs when the
// epilogue mode is 1, see nsp_set_gemm_epilogue).  Why a second epilogue: in gemm_tcgen05.cu every epilogue thread
// owns one output row and writes it with 16-byte vector stores, i.e. each warp-level store instruction touches 32
// different 128-byte lines (32 L1 wavefronts, ~2 cycles each).  A 128x128 fp32 tile costs ~4096 wavefronts = ~8000
// cycles of LSU time against 2048 cycles of tensor-core time at K = 512: the K <= 1024 GEMMs of the encoder are
// epilogue-paced (NOTES.md).  Here the accumulators go TMEM -> registers -> a swizzled 32 x 128-byte staging tile in
// shared memory (conflict-free 16-byte st.shared) -> one cp.async.bulk.tensor store per 32x32 sub-tile; the
// residual operand comes in the same way (TMA load into the staging tile while the mainloop of the tile is still
// running, result written back in place).  TMA clips rows >= M / columns >= N, so there is no tail handling.
//
// Staging layout per epilogue warp: fp32 sub-tile = 32 rows x 128 B, SWIZZLE_128B (16-byte unit j of row r lives at
// unit j ^ (r & 7)); bf16 sub-tile = 32 rows x 64 B, SWIZZLE_64B (unit j of row r at j ^ ((r >> 1) & 3)).
#include "tc_common.cuh"

namespace nsp {

bool get_tma_encode(void** fn);   // gemm_tcgen05.cu

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// row-major [rows, cols] matrix, box = [box_rows, box_cols]; box_cols * elem_bytes must equal the swizzle span
bool encode_tmap_2d(CUtensorMap* out, const void* base, bool is_bf16, uint64_t rows, uint64_t cols, uint64_t ld,
               uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swz, const char* what) {
    void* fnp = nullptr;
    if (!get_tma_encode(&fnp)) return false;
    const int es = is_bf16 ? 2 : 4;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * (uint64_t)es};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fnp)(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                                      const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(%s) failed (%d) rows=%llu cols=%llu ld=%llu", what, (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
        return false;
    }
    return true;
}

namespace {

long long g_ts_launches = 0;         // how many GEMMs took this kernel (tests check that the envelope logic routes here)
long long g_pair_launches = 0;       // ... of which as CTA pairs

constexpr int BM = 128;
constexpr int KBYTES = 128;
constexpr int BK = 64;               // bf16 elements per 128-byte stage row
constexpr int UK = 16;               // K per tcgen05.mma (bf16)
constexpr int NEPI_WARPS = 8;
constexpr int NTHREADS = 64 + 32 * NEPI_WARPS;
constexpr int CW = 32;               // accumulator columns per sub-tile

struct TsMaps {
    CUtensorMap a, b, out, pre, res;
};

struct TsArgs {
    int M, N, K;            // N = rows of W (2 * output columns for GLU)
    const float* bias;      // [N] or null
    float alpha;
    int has_pre;            // save the pre-activation through maps.pre (bf16 [M, N])
};

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SWISH = 2, ACT_GELU = 3, ACT_GELU_TANH = 4 };

__device__ __forceinline__ float sigmoid_fast(float x) {     // 0.5 + 0.5 * tanh(x / 2): one MUFU
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
}

// ---- CTA-pair (cta_group::2) plumbing: PTX forms as used by cute/arch/{copy_sm100_tma,mma_sm100_umma}.hpp and
// cutlass/arch/barrier.h for 2-SM kernels ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {            // every thread of both CTAs
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {     // shared::cluster address in CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
// tile into the executing CTA's shared memory, completion bytes onto a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(tc::smem_u32(smem_dst)), "l"(m), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_holder) {   // one full warp in EACH CTA, same warp index
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(tc::smem_u32(smem_holder)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(NCOLS) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 per CTA] * B[N rows: N/2 per CTA]^T, issued by the leader CTA only
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// one arrival on the barrier at the same shared-memory offset in both CTAs when the leader's MMAs retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(tc::smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// bounded mbarrier wait (tc_common.cuh: bare spin + one counter, the clock is read once per 4096 failed polls)
__device__ __forceinline__ void bwait(uint64_t* bar, uint32_t parity, int) { tc::mbar_wait(bar, parity); }

// TWO = CTA pair (cluster of 2, cta_group::2): the pair computes a 256 x BN tile; each CTA stages its own 128 rows of A and
// BN/2 rows of W per k-block (half the L2->SMEM operand traffic per FLOP of the single-CTA tile), the leader CTA issues
// the M = 256 MMAs for both, and each CTA runs the epilogue of its own 128 accumulator rows.
template <int BN, int STAGES, int ACT, bool GLU, bool RES, bool OUTBF16, bool TWO>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_ts_kernel(const __grid_constant__ TsMaps maps, const TsArgs g) {
    pdl_launch_dependents();      // PDL: the next kernel may start its prologue; ours overlaps the previous kernel's tail
    static_assert(!(TWO && GLU), "GLU is single-CTA only");
    constexpr int A_BYTES = BM * KBYTES;
    constexpr int B_BYTES = (TWO ? BN / 2 : BN) * KBYTES;
    constexpr int TILE_M = TWO ? 2 * BM : BM;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = 2 * BN;
    constexpr int BN_OUT = GLU ? BN / 2 : BN;
    constexpr int COLS_PER_WARP = BN_OUT / 2;
    constexpr int NCH = COLS_PER_WARP / CW;
    constexpr int SLOT = OUTBF16 ? 2048 : 4096;               // one output sub-tile
    constexpr int STG_WARP = RES ? NCH * 4096 : 2 * SLOT;      // residual: the warp's whole region; else two slots in rotation
    static_assert(COLS_PER_WARP % CW == 0, "tile width");
    static_assert(!RES || !OUTBF16, "residual epilogue is fp32");
    static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns");

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* stg_all = smem + STAGES * STAGE_BYTES;                          // 1024-byte aligned (stage sizes are multiples)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_all + NEPI_WARPS * STG_WARP);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;    // [2]
    uint64_t* tempty_bar = tfull_bar + 2;        // [2]
    uint64_t* res_bar = tempty_bar + 2;          // [NEPI_WARPS]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(res_bar + NEPI_WARPS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nout = GLU ? g.N / 2 : g.N;
    const int m_tiles = (g.M + TILE_M - 1) / TILE_M;
    const int n_tiles = (nout + BN_OUT - 1) / BN_OUT;
    const int num_tiles = m_tiles * n_tiles;
    const int kblocks = (g.K + BK - 1) / BK;
    const uint32_t rank = TWO ? cluster_ctarank() : 0u;              // 0 = leader of the pair
    const int tile0 = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&maps.a); tc::tma_prefetch_desc(&maps.b); tc::tma_prefetch_desc(&maps.out);
        if (g.has_pre) tc::tma_prefetch_desc(&maps.pre);
        if constexpr (RES) tc::tma_prefetch_desc(&maps.res);
    }
    if (warp == 1 && lane == 0) {
        // pair: the full / accumulator-empty barriers that matter are the leader's (one arrival per CTA's producer, one per
        // epilogue warp of both CTAs); stage-empty / accumulator-full are per CTA, signalled by the leader's multicast commit
        for (int i = 0; i < STAGES; ++i) { tc::mbar_init(&full_bar[i], TWO ? 2 : 1); tc::mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull_bar[i], 1); tc::mbar_init(&tempty_bar[i], (TWO ? 2 : 1) * NEPI_WARPS); }
        for (int i = 0; i < NEPI_WARPS; ++i) tc::mbar_init(&res_bar[i], 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (TWO) tmem_alloc_pair<TMEM_COLS>(tmem_holder);
        else tc::tmem_alloc<TMEM_COLS>(tmem_holder);
    }
    tc::tc_fence_before();
    if constexpr (TWO) cluster_sync_all();       // barrier inits and the allocation of BOTH CTAs are visible before any traffic
    else __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();                    // the previous grid is complete: operands / residuals / outputs may be touched from here

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (tc::elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
                for (int kb = 0; kb < kblocks; ++kb) {
                    bwait(&empty_bar[stage], phase ^ 1, 0);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    if constexpr (TWO) {
                        const uint32_t lead_bar = mapa_rank(tc::smem_u32(&full_bar[stage]), 0);
                        if (rank == 0) tc::mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);   // bytes of both CTAs
                        tma_load_2d_pair(sa, &maps.a, lead_bar, kb * BK, m_blk * TILE_M + (int)rank * BM);
                        tma_load_2d_pair(sb, &maps.b, lead_bar, kb * BK, n_blk * BN + (int)rank * (BN / 2));
                        if (rank != 0) mbar_arrive_cluster(lead_bar);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    tc::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                    tc::tma_load_2d(sa, &maps.a, &full_bar[stage], kb * BK, m_blk * BM);
                    if constexpr (!GLU) {
                        tc::tma_load_2d(sb, &maps.b, &full_bar[stage], kb * BK, n_blk * BN);
                    } else {   // value rows then gate rows of W land in one BN-row smem tile
                        tc::tma_load_2d(sb, &maps.b, &full_bar[stage], kb * BK, n_blk * (BN / 2));
                        tc::tma_load_2d(sb + B_BYTES / 2, &maps.b, &full_bar[stage], kb * BK, g.N / 2 + n_blk * (BN / 2));
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (single thread) =====================
        if (rank == 0 && tc::elect_one()) {
            constexpr uint32_t idesc = tc::make_idesc(1u, TILE_M, BN);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                bwait(&tempty_bar[as], aphase ^ 1, 1);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                uint32_t accum = 0;
                for (int kb = 0; kb < kblocks; ++kb) {
                    bwait(&full_bar[stage], phase, 2);
                    tc::tc_fence_after();
                    const uint32_t sa = tc::smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t adesc = tc::make_smem_desc_sw128(sa);
                    const uint64_t bdesc = tc::make_smem_desc_sw128(sa + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k) {
                        const uint64_t koff = (uint64_t)((k * 32) >> 4);     // +32 bytes of K inside the swizzle atom
                        if constexpr (TWO) umma_f16_pair(d_tmem, adesc + koff, bdesc + koff, idesc, accum);
                        else tc::umma_f16(d_tmem, adesc + koff, bdesc + koff, idesc, accum);
                        accum = 1;
                    }
                    if constexpr (TWO) umma_commit_pair(&empty_bar[stage]);
                    else tc::umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if constexpr (TWO) umma_commit_pair(&tfull_bar[as]);
                else tc::umma_commit(&tfull_bar[as]);
            }
        }
    } else {
        // ===================== epilogue warps (2..9) =====================
        const int ew = warp - 2;
        const int q = warp & 3;                                 // TMEM lane quadrant this warp may access
        const int half = ew >> 2;                               // which half of the tile's columns
        uint8_t* stg = stg_all + ew * STG_WARP;
        uint64_t* rbar = &res_bar[ew];
        const bool has_bias = g.bias != nullptr;
        const bool bias_vec = (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0;
        uint32_t rphase = 0;
        int slot = 0;
        int it = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
            const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int row0 = m_blk * TILE_M + (int)rank * BM + q * 32;   // first row of this warp's sub-tiles
            const int col0 = n_blk * BN_OUT + half * COLS_PER_WARP;
            const bool live = row0 < g.M && col0 < nout;        // warp-uniform: anything to write at all
            if constexpr (RES) {
                // the residual sub-tiles do not depend on the accumulator: fetch them under the mainloop of this tile
                if (live && tc::elect_sync()) {
                    tc::bulk_wait_read<0>();                        // stores of the previous tile have left the staging tiles
                    tc::mbar_arrive_expect_tx(rbar, NCH * 4096);
#pragma unroll
                    for (int ci = 0; ci < NCH; ++ci) tc::tma_load_2d(stg + ci * 4096, &maps.res, rbar, col0 + ci * CW, row0);
                }
            }
            tc::mbar_wait_epi(&tfull_bar[as], aphase);
            tc::tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
            if constexpr (RES) {
                if (live) { bwait(rbar, rphase, 4); rphase ^= 1; }
            }
            uint32_t rbuf[2][32];
            tc::tmem_ld_32x32(t_row + (uint32_t)(half * COLS_PER_WARP), rbuf[0]);
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                const int cbase = col0 + ci * CW;
                const int tcol = half * COLS_PER_WARP + ci * CW;     // column inside the accumulator (value part)
                float v[32];
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(rbuf[ci & 1][j]);
                if (ci + 1 < NCH) tc::tmem_ld_32x32(t_row + (uint32_t)(tcol + CW), rbuf[(ci + 1) & 1]);
                float gv[32];
                if constexpr (GLU) {                                 // gate half of the accumulator (warp-collective load)
                    uint32_t r2[32];
                    tc::tmem_ld_32x32(t_row + (uint32_t)(BN / 2 + tcol), r2);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < CW; ++j) gv[j] = __uint_as_float(r2[j]);
                }
                if (!(live && cbase < nout)) continue;               // warp-uniform; nout % 32 == 0 (host check)
                if (has_bias) {
                    if (bias_vec) {
#pragma unroll
                        for (int j = 0; j < CW; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + cbase + j));
                            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < CW; ++j) v[j] += __ldg(g.bias + cbase + j);
                    }
                    if constexpr (GLU) {
                        if (bias_vec && (nout % 4 == 0)) {
#pragma unroll
                            for (int j = 0; j < CW; j += 4) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + nout + cbase + j));
                                gv[j] += b4.x; gv[j + 1] += b4.y; gv[j + 2] += b4.z; gv[j + 3] += b4.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < CW; ++j) gv[j] += __ldg(g.bias + nout + cbase + j);
                        }
                    }
                }
                if (g.has_pre) {                                     // pre-activation(s), bf16, through the slot rotation
                    if constexpr (!RES) {
                        if (tc::elect_sync()) tc::bulk_wait_read<1>();
                        __syncwarp();
                        tc::st_row_bf16(stg + slot * SLOT, lane, v);
                        tc::fence_proxy_async_smem();
                        __syncwarp();
                        if (tc::elect_one()) { tc::tma_store_2d(&maps.pre, stg + slot * SLOT, cbase, row0); tc::bulk_commit(); }
                        slot ^= 1;
                        if constexpr (GLU) {
                            if (tc::elect_sync()) tc::bulk_wait_read<1>();
                            __syncwarp();
                            tc::st_row_bf16(stg + slot * SLOT, lane, gv);
                            tc::fence_proxy_async_smem();
                            __syncwarp();
                            if (tc::elect_one()) { tc::tma_store_2d(&maps.pre, stg + slot * SLOT, nout + cbase, row0); tc::bulk_commit(); }
                            slot ^= 1;
                        }
                    }
                }
                if constexpr (GLU) {
#pragma unroll
                    for (int j = 0; j < CW; ++j) v[j] *= sigmoid_fast(gv[j]);
                } else if constexpr (ACT == ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < CW; ++j) v[j] = fmaxf(v[j], 0.f);
                } else if constexpr (ACT == ACT_SWISH) {
#pragma unroll
                    for (int j = 0; j < CW; ++j) v[j] *= sigmoid_fast(v[j]);
                } else if constexpr (ACT == ACT_GELU) {              // F.gelu (modules/gelu.py:17-21): x * Phi(x)
#pragma unroll
                    for (int j = 0; j < CW; ++j) v[j] = 0.5f * v[j] * (1.f + erff(v[j] * 0.70710678118654752f));
                } else if constexpr (ACT == ACT_GELU_TANH) {         // tanh approximation (modules/gelu.py:11-14)
#pragma unroll
                    for (int j = 0; j < CW; ++j) {
                        const float x = v[j];
                        v[j] = 0.5f * x * (1.f + tanhf(0.79788456080286536f * (x + 0.044715f * x * x * x)));
                    }
                }
                if constexpr (RES) {
                    uint8_t* tile_s = stg + ci * 4096;
                    float rr[32];
                    tc::ld_row_f32(tile_s, lane, rr);
#pragma unroll
                    for (int j = 0; j < CW; ++j) v[j] = fmaf(g.alpha, v[j], rr[j]);
                    tc::st_row_f32(tile_s, lane, v);                     // in place: a thread only touches its own row
                    tc::fence_proxy_async_smem();
                    __syncwarp();
                    if (tc::elect_one()) { tc::tma_store_2d(&maps.out, tile_s, cbase, row0); tc::bulk_commit(); }
                } else {
                    if (g.alpha != 1.f) {
#pragma unroll
                        for (int j = 0; j < CW; ++j) v[j] *= g.alpha;
                    }
                    if (tc::elect_sync()) tc::bulk_wait_read<1>();              // the store that used this slot two stores ago is done
                    __syncwarp();
                    uint8_t* tile_s = stg + slot * SLOT;
                    if constexpr (OUTBF16) tc::st_row_bf16(tile_s, lane, v);
                    else tc::st_row_f32(tile_s, lane, v);
                    tc::fence_proxy_async_smem();
                    __syncwarp();
                    if (tc::elect_one()) { tc::tma_store_2d(&maps.out, tile_s, cbase, row0); tc::bulk_commit(); }
                    slot ^= 1;
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (TWO) mbar_arrive_cluster(mapa_rank(tc::smem_u32(&tempty_bar[as]), 0));   // the leader's MMA issuer waits
                else tc::mbar_arrive(&tempty_bar[as]);
            }
        }
        if (tc::elect_sync()) tc::bulk_wait_read<0>();                          // staging tiles must outlive the bulk stores reading them
    }

    tc::tc_fence_before();
    if constexpr (TWO) cluster_sync_all();       // neither CTA may retire (or free TMEM) while the pair's MMAs / arrivals can touch it
    else __syncthreads();
    if (warp == 2) {
        tc::tc_fence_after();
        if constexpr (TWO) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
        else tc::tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

template <int BN, bool GLU, bool RES, bool OUTBF16, bool TWO>
constexpr int ts_staging_bytes() {
    return NEPI_WARPS * (RES ? ((GLU ? BN / 2 : BN) / 2 / CW) * 4096 : 2 * (OUTBF16 ? 2048 : 4096));
}
template <int BN, bool GLU, bool RES, bool OUTBF16, bool TWO>
constexpr int ts_stages() {
    constexpr int stage = BM * KBYTES + (TWO ? BN / 2 : BN) * KBYTES;
    constexpr int n = (224 * 1024 - ts_staging_bytes<BN, GLU, RES, OUTBF16, TWO>()) / stage;
    return n > 8 ? 8 : n;
}

template <int BN, int ACT, bool GLU, bool RES, bool OUTBF16, bool TWO>
nsp_status launch_ts(const TsMaps& maps, const TsArgs& g, cudaStream_t st) {
    constexpr int STAGES = ts_stages<BN, GLU, RES, OUTBF16, TWO>();
    constexpr size_t smem = (size_t)STAGES * (BM * KBYTES + (TWO ? BN / 2 : BN) * KBYTES) + ts_staging_bytes<BN, GLU, RES, OUTBF16, TWO>()
                            + 1024 /*align*/ + 512 /*barriers*/;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static_assert(STAGES >= 3, "pipeline depth");
    auto kern = gemm_ts_kernel<BN, STAGES, ACT, GLU, RES, OUTBF16, TWO>;
    static bool attr_set = false;
    if (!attr_set) {
        NSP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int nout = GLU ? g.N / 2 : g.N;
    const int bn_out = GLU ? BN / 2 : BN;
    const int tiles = ceil_div(g.M, TWO ? 2 * BM : BM) * ceil_div(nout, bn_out);
    if constexpr (TWO) {
        const int pairs = num_sms() / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * (tiles < pairs ? tiles : pairs));
        cfg.blockDim = dim3(NTHREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl_enabled() ? 2 : 1;
        NSP_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, maps, g));
    } else {
        const int grid = tiles < num_sms() ? tiles : num_sms();
        launch_k(kern, dim3(grid), dim3(NTHREADS), smem, st, maps, g);
        NSP_LAUNCH_OK();
    }
    ++g_ts_launches;
    if (TWO) ++g_pair_launches;
    return NSP_OK;
}

template <int BN, bool TWO>
nsp_status dispatch_ts(const TsMaps& maps, const TsArgs& g, int glu, int act, bool res, int out_bf16, cudaStream_t st) {
    if (glu) {
        if constexpr (BN == 128 && !TWO) {
            return out_bf16 ? launch_ts<128, ACT_NONE, true, false, true, false>(maps, g, st)
                            : launch_ts<128, ACT_NONE, true, false, false, false>(maps, g, st);
        } else {
            set_error("gemm(tma epilogue): GLU needs the single-CTA 128-wide tile"); return NSP_ERR_INVALID;
        }
    }
    if (res) {
        if constexpr (BN == 128) return launch_ts<128, ACT_NONE, false, true, false, TWO>(maps, g, st);
        else { set_error("gemm(tma epilogue): residual needs the 128-wide tile"); return NSP_ERR_INVALID; }
    }
    switch (act) {
        case ACT_NONE: return out_bf16 ? launch_ts<BN, ACT_NONE, false, false, true, TWO>(maps, g, st)
                                       : launch_ts<BN, ACT_NONE, false, false, false, TWO>(maps, g, st);
        case ACT_RELU: return out_bf16 ? launch_ts<BN, ACT_RELU, false, false, true, TWO>(maps, g, st)
                                       : launch_ts<BN, ACT_RELU, false, false, false, TWO>(maps, g, st);
        case ACT_SWISH: return out_bf16 ? launch_ts<BN, ACT_SWISH, false, false, true, TWO>(maps, g, st)
                                        : launch_ts<BN, ACT_SWISH, false, false, false, TWO>(maps, g, st);
        case ACT_GELU: return out_bf16 ? launch_ts<BN, ACT_GELU, false, false, true, TWO>(maps, g, st)
                                       : launch_ts<BN, ACT_GELU, false, false, false, TWO>(maps, g, st);
        case ACT_GELU_TANH: return out_bf16 ? launch_ts<BN, ACT_GELU_TANH, false, false, true, TWO>(maps, g, st)
                                            : launch_ts<BN, ACT_GELU_TANH, false, false, false, TWO>(maps, g, st);
    }
    set_error("gemm(tma epilogue): act=%d", act);
    return NSP_ERR_INVALID;
}

bool aligned16(const void* p, int64_t ld, int es) { return ((uintptr_t)p % 16 == 0) && ((ld * es) % 16 == 0); }

}  // namespace

long long gemm_ts_launch_count() { return g_ts_launches; }
long long gemm_pair_launch_count() { return g_pair_launches; }

// Called by gemm_dispatch (gemm_tcgen05.cu) for bf16 operands when the epilogue mode is 1 or 2.  bn1 = the tile width the
// single-CTA heuristic picked.  *handled = false means "outside this kernel's envelope" (the caller runs the direct-store
// kernel); otherwise the return value is final.
nsp_status gemm_ts_dispatch(int mode, const void* a, int64_t lda, const void* w, int64_t ldw, int M, int N, int K, int glu,
                            int act, const float* bias, const float* residual, int64_t ldr, float alpha, void* out, int64_t ldo,
                            int out_bf16, void* pre, int64_t ldpre, int bn1, cudaStream_t st, bool* handled) {
    *handled = false;
    const int nout = glu ? N / 2 : N;
    const bool res = residual != nullptr;
    if (nout % CW != 0) return NSP_OK;
    if (res && (out_bf16 || pre || act != ACT_NONE || glu)) return NSP_OK;
    if (!aligned16(out, ldo, out_bf16 ? 2 : 4)) return NSP_OK;
    if (pre && !aligned16(pre, ldpre, 2)) return NSP_OK;
    if (res && !aligned16(residual, ldr, 4)) return NSP_OK;
    if (bias && ((uintptr_t)bias % 4 != 0)) return NSP_OK;
    // CTA pairs (mode 2): 256 x {256,128} tiles when they still fill most of the 74 pairs
    bool two = false;
    int BN = bn1;
    if (mode == 2 && !glu) {
        const int bn2 = (!res && nout % 256 == 0) ? 256 : 128;
        const int64_t tiles2 = (int64_t)ceil_div(M, 2 * BM) * ceil_div(nout, bn2);
        if (tiles2 * 5 >= (int64_t)(num_sms() / 2) * 4) { two = true; BN = bn2; }
    }
    if (!two) {
        if (BN != 128 && BN != 256) return NSP_OK;                   // small problems stay on the direct-store kernel
        if (res && BN != 128) return NSP_OK;
        if (glu && BN != 128) return NSP_OK;
    }
    *handled = true;
    TsMaps maps;
    TsArgs g;
    g.M = M; g.N = N; g.K = K; g.bias = bias; g.alpha = alpha; g.has_pre = pre != nullptr;
    const uint32_t box_b = two ? (uint32_t)(BN / 2) : (glu ? (uint32_t)(BN / 2) : (uint32_t)BN);
    if (!make_tmap_2d(&maps.a, a, 2, true, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM)) return NSP_ERR_INVALID;
    if (!make_tmap_2d(&maps.b, w, 2, true, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, box_b)) return NSP_ERR_INVALID;
    if (!encode_tmap_2d(&maps.out, out, out_bf16 != 0, (uint64_t)M, (uint64_t)nout, (uint64_t)ldo, 32, CW,
                   out_bf16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, "out")) return NSP_ERR_INVALID;
    maps.pre = maps.out;
    maps.res = maps.out;
    if (pre && !encode_tmap_2d(&maps.pre, pre, true, (uint64_t)M, (uint64_t)N, (uint64_t)ldpre, 32, CW, CU_TENSOR_MAP_SWIZZLE_64B, "pre"))
        return NSP_ERR_INVALID;
    if (res && !encode_tmap_2d(&maps.res, residual, false, (uint64_t)M, (uint64_t)nout, (uint64_t)ldr, 32, CW,
                          CU_TENSOR_MAP_SWIZZLE_128B, "residual")) return NSP_ERR_INVALID;
    if (two) return BN == 256 ? dispatch_ts<256, true>(maps, g, glu, act, res, out_bf16, st)
                              : dispatch_ts<128, true>(maps, g, glu, act, res, out_bf16, st);
    return BN == 256 ? dispatch_ts<256, false>(maps, g, glu, act, res, out_bf16, st)
                     : dispatch_ts<128, false>(maps, g, glu, act, res, out_bf16, st);
}

}  // namespace nsp
