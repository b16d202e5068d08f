// Persistent, warp-specialised tcgen05 GEMM for sm_100a with fused epilogues:
//     out = residual + alpha * act(A[M,K] * W[N,K]^T + bias)        (act: none / relu / swish / GLU)
//
// Replaces every nn.Linear / 1x1 Conv1d on the encoder path of the reference:
//   PositionwiseFeedForward w_1/w_2   neural_sp/models/modules/positionwise_feed_forward.py:47-48,89
//   RelMHA w_key/w_value/w_query/w_out/w_pos   .../relative_multihead_attention.py:57-60,169-176,217
//   ConformerConvBlock pointwise_conv1(+GLU)/pointwise_conv2   .../conformer_convolution.py:44-69,110-126
//   ConvEncoder.bridge  encoders/conv.py:87,193 ; CTC output head  decoders/ctc.py:81-91,124
//
// Structure (one CTA per SM, 192 threads):
//   warp 0   TMA producer: cp.async.bulk.tensor 128B-swizzled A/W tiles -> STAGES-deep smem ring (mbarrier full/empty)
//   warp 1   MMA issuer : one thread issues tcgen05.mma (M=128, N=BN, K=16 bf16 / 8 tf32), accumulators in TMEM,
//            double-buffered (2 x BN columns) so the epilogue of tile i overlaps the mainloop of tile i+1
//   warps 2-5 epilogue : tcgen05.ld TMEM -> registers -> bias/activation/residual -> 128-bit global stores
// Operands are K-major (row-major activations and nn.Linear weights as stored), so no transposes are needed.
// Parity mode ("fp32") runs three tf32 segments over pre-split hi/lo operands into the same accumulator
// (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo), which restores ~fp32 accuracy on the tensor cores.
#include "tc_common.cuh"

namespace nsp {

static bool get_encode_fn(void** fn) {
    static void* cached = nullptr;
    if (!cached) {
        cudaDriverEntryPointQueryResult q;
        void* f = nullptr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !f) {
            set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
            (void)cudaGetLastError();
            return false;
        }
        cached = f;
    }
    *fn = cached;
    return true;
}

bool get_tma_encode(void** fn) { return get_encode_fn(fn); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, bool is_bf16, uint64_t rows, uint64_t cols,
                  uint64_t ld, uint32_t box_rows) {
    void* fnp = nullptr;
    if (!get_encode_fn(&fnp)) return false;
    EncodeTiledFn fn = (EncodeTiledFn)fnp;
    if (((uintptr_t)base) % 16 != 0 || (ld * elem_bytes) % 16 != 0) {
        set_error("TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch (ptr=%p ld=%llu elem=%d)",
                  base, (unsigned long long)ld, elem_bytes);
        return false;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * (uint64_t)elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                    const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box_rows=%u", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
        return false;
    }
    return true;
}

namespace {

constexpr int BM = 128;
constexpr int KBYTES = 128;          // bytes of K per stage row (one 128B swizzle atom)
constexpr int NEPI_WARPS = 8;        // two warps per TMEM lane quadrant, each takes half of the tile's columns
constexpr int NTHREADS = 64 + 32 * NEPI_WARPS;
constexpr int MAX_SEG = 3;

struct GemmMaps {
    CUtensorMap a[MAX_SEG];
    CUtensorMap b[MAX_SEG];
};

struct GemmArgs {
    int M, N, K;            // N = rows of W (2*Nout for GLU)
    int nseg;
    const float* bias;      // [N] or null
    const float* residual;  // [M, ldr] fp32 or null
    int64_t ldr;
    float alpha;
    void* out;              // [M, ldo] fp32 or bf16
    int64_t ldo;
    void* out2;             // optional second output (bf16 copy of an fp32 result), or null
    int64_t ldo2;
    void* pre;              // optional: pre-activation (acc + bias) saved for the backward pass, operand dtype
    int64_t ldpre;          //           [M, N] (GLU: value columns [0,N/2), gate columns [N/2,N)), or null
};

// epilogue flavours (compile-time: the epilogue is instruction-bound, see profiles/r01_gemm_epilogue.md)
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SWISH = 2, ACT_GELU = 3, ACT_GELU_TANH = 4 };

template <bool FAST>
__device__ __forceinline__ float sigmoid_f(float x) {
    if constexpr (FAST) {           // 0.5 + 0.5 * tanh(x/2): one MUFU instead of EX2 + RCP (bf16 operand mode)
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
        return fmaf(0.5f, t, 0.5f);
    } else {
        return __fdividef(1.f, 1.f + __expf(-x));
    }
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&p);
}

// store CW consecutive values of one row in the operand dtype (bf16 or fp32); nvalid = columns inside the matrix
template <bool BF16, int CW>
__device__ __forceinline__ void store_pre(void* base, int64_t ld, int row, int col, const float* v, int nvalid) {
    if constexpr (BF16) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(base) + (int64_t)row * ld + col;
        if (nvalid >= CW && (ld % 8 == 0) && (col % 8 == 0)) {
#pragma unroll
            for (int j = 0; j < CW; j += 8) {
                uint4 pk;
                pk.x = pack_bf16(v[j], v[j + 1]); pk.y = pack_bf16(v[j + 2], v[j + 3]);
                pk.z = pack_bf16(v[j + 4], v[j + 5]); pk.w = pack_bf16(v[j + 6], v[j + 7]);
                *reinterpret_cast<uint4*>(o + j) = pk;
            }
        } else {
#pragma unroll
            for (int j = 0; j < CW; ++j) if (j < nvalid) o[j] = __float2bfloat16_rn(v[j]);
        }
    } else {
        float* o = reinterpret_cast<float*>(base) + (int64_t)row * ld + col;
        if (nvalid >= CW && (ld % 4 == 0) && (col % 4 == 0)) {
#pragma unroll
            for (int j = 0; j < CW; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < CW; ++j) if (j < nvalid) o[j] = v[j];
        }
    }
}

template <typename TIn, int BN, int STAGES, int ACT, bool GLU, bool RES, bool OUTBF16>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_kernel(const __grid_constant__ GemmMaps maps, const GemmArgs g) {
    pdl_launch_dependents();      // PDL: the next kernel may start its prologue; ours overlaps the previous kernel's tail
    constexpr bool kBF16 = sizeof(TIn) == 2;
    constexpr int BK = KBYTES / (int)sizeof(TIn);        // 64 (bf16) or 32 (tf32)
    constexpr int UK = 32 / (int)sizeof(TIn);            // K per tcgen05.mma: 16 (bf16) or 8 (tf32)
    constexpr int A_BYTES = BM * KBYTES;                  // 16 KiB
    constexpr int B_BYTES = BN * KBYTES;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = 2 * BN;                     // two accumulator stages
    constexpr int BN_OUT = GLU ? BN / 2 : BN;             // output columns per tile
    constexpr int COLS_PER_WARP = BN_OUT / 2;             // each quadrant is shared by two warps
    static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns must be a power of two");
    static_assert(COLS_PER_WARP % 32 == 0 || COLS_PER_WARP == 16, "unsupported tile width");

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;    // [2]
    uint64_t* tempty_bar = tfull_bar + 2;        // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nout = GLU ? g.N / 2 : g.N;                 // output columns
    const int m_tiles = (g.M + BM - 1) / BM;
    const int n_tiles = (nout + BN_OUT - 1) / BN_OUT;
    const int num_tiles = m_tiles * n_tiles;
    const int kblocks = (g.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < g.nseg; ++s) { tc::tma_prefetch_desc(&maps.a[s]); tc::tma_prefetch_desc(&maps.b[s]); }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull_bar[i], 1); tc::mbar_init(&tempty_bar[i], NEPI_WARPS); }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<TMEM_COLS>(tmem_holder);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();                    // the previous grid is complete: operands / residuals / outputs may be touched from here

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (tc::elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
                for (int s = 0; s < g.nseg; ++s) {
                    for (int kb = 0; kb < kblocks; ++kb) {
                        tc::mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * STAGE_BYTES;
                        uint8_t* sb = sa + A_BYTES;
                        tc::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                        tc::tma_load_2d(sa, &maps.a[s], &full_bar[stage], kb * BK, m_blk * BM);
                        if constexpr (!GLU) {
                            tc::tma_load_2d(sb, &maps.b[s], &full_bar[stage], kb * BK, n_blk * BN);
                        } else {   // value rows then gate rows of W land in one BN-row smem tile
                            tc::tma_load_2d(sb, &maps.b[s], &full_bar[stage], kb * BK, n_blk * (BN / 2));
                            tc::tma_load_2d(sb + B_BYTES / 2, &maps.b[s], &full_bar[stage], kb * BK,
                                            g.N / 2 + n_blk * (BN / 2));
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (single thread) =====================
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::make_idesc(kBF16 ? 1u : 2u, BM, BN);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                tc::mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                uint32_t accum = 0;
                for (int s = 0; s < g.nseg; ++s) {
                    for (int kb = 0; kb < kblocks; ++kb) {
                        tc::mbar_wait(&full_bar[stage], phase);
                        tc::tc_fence_after();
                        const uint32_t sa = tc::smem_u32(smem + stage * STAGE_BYTES);
                        const uint64_t adesc = tc::make_smem_desc_sw128(sa);
                        const uint64_t bdesc = tc::make_smem_desc_sw128(sa + A_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k) {
                            // advance K inside the 128B swizzle atom: +32 bytes per MMA (>>4 in the descriptor)
                            const uint64_t koff = (uint64_t)((k * 32) >> 4);
                            if constexpr (kBF16) tc::umma_f16(d_tmem, adesc + koff, bdesc + koff, idesc, accum);
                            else tc::umma_tf32(d_tmem, adesc + koff, bdesc + koff, idesc, accum);
                            accum = 1;
                        }
                        tc::umma_commit(&empty_bar[stage]);     // smem slot reusable once these MMAs retire
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                tc::umma_commit(&tfull_bar[as]);                // accumulator complete -> epilogue
            }
        }
    } else {
        // ===================== epilogue warps (2..9) =====================
        const int q = warp & 3;                                 // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;                       // which half of the tile's columns
        const bool has_bias = g.bias != nullptr;
        const bool bias_vec = (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            tc::mbar_wait_epi(&tfull_bar[as], aphase);
            tc::tc_fence_after();
            const int row = m_blk * BM + q * 32 + lane;
            const bool row_ok = row < g.M;
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
            const int col0 = n_blk * BN_OUT + half * COLS_PER_WARP;
            constexpr int CW = COLS_PER_WARP >= 32 ? 32 : 16;   // columns per chunk
            constexpr int NCH = COLS_PER_WARP / CW;
            // software pipeline over the chunks: the TMEM load of chunk c+1 is in flight while chunk c is processed
            uint32_t rbuf[2][32];
            if constexpr (CW == 32) tc::tmem_ld_32x32(t_row + (uint32_t)(half * COLS_PER_WARP), rbuf[0]);
            else tc::tmem_ld_32x16(t_row + (uint32_t)(half * COLS_PER_WARP), rbuf[0]);
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                const int c = ci * CW;
                float v[32];
                const int cbase = col0 + c;
                const int tcol = half * COLS_PER_WARP + c;      // column inside the accumulator (value part)
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(rbuf[ci & 1][j]);
                if (ci + 1 < NCH) {
                    if constexpr (CW == 32) tc::tmem_ld_32x32(t_row + (uint32_t)(tcol + CW), rbuf[(ci + 1) & 1]);
                    else tc::tmem_ld_32x16(t_row + (uint32_t)(tcol + CW), rbuf[(ci + 1) & 1]);
                }
                const bool full = (cbase + CW <= nout);
                if (has_bias) {
                    if (full && bias_vec) {
#pragma unroll
                        for (int j = 0; j < CW; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + cbase + j));
                            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < CW; ++j) if (cbase + j < nout) v[j] += __ldg(g.bias + cbase + j);
                    }
                }
                if (g.pre && row_ok) store_pre<kBF16, CW>(g.pre, g.ldpre, row, cbase, v, nout - cbase);
                if constexpr (GLU) {
                    uint32_t r2[32];
                    if constexpr (CW == 32) tc::tmem_ld_32x32(t_row + (uint32_t)(BN / 2 + tcol), r2);
                    else tc::tmem_ld_32x16(t_row + (uint32_t)(BN / 2 + tcol), r2);
                    tc::tmem_ld_wait();
                    float gv[CW];
#pragma unroll
                    for (int j = 0; j < CW; ++j) {
                        float gt = __uint_as_float(r2[j]);
                        if (has_bias && cbase + j < nout) gt += __ldg(g.bias + nout + cbase + j);
                        gv[j] = gt;
                        v[j] *= sigmoid_f<kBF16>(gt);
                    }
                    if (g.pre && row_ok) store_pre<kBF16, CW>(g.pre, g.ldpre, row, nout + cbase, gv, nout - cbase);
                } else if constexpr (ACT == ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < CW; ++j) v[j] = fmaxf(v[j], 0.f);
                } else if constexpr (ACT == ACT_SWISH) {
#pragma unroll
                    for (int j = 0; j < CW; ++j) v[j] *= sigmoid_f<kBF16>(v[j]);
                } else if constexpr (ACT == ACT_GELU) {          // F.gelu (modules/gelu.py:17-21): x * Phi(x)
#pragma unroll
                    for (int j = 0; j < CW; ++j) v[j] = 0.5f * v[j] * (1.f + erff(v[j] * 0.70710678118654752f));
                } else if constexpr (ACT == ACT_GELU_TANH) {     // tanh approximation (modules/gelu.py:11-14)
#pragma unroll
                    for (int j = 0; j < CW; ++j) {
                        const float x = v[j];
                        v[j] = 0.5f * x * (1.f + tanhf(0.79788456080286536f * (x + 0.044715f * x * x * x)));
                    }
                }
                if (row_ok) {
                    if constexpr (RES) {
                        const float* rr = g.residual + (int64_t)row * g.ldr + cbase;
                        if (full && (g.ldr % 4 == 0)) {
#pragma unroll
                            for (int j = 0; j < CW; j += 4) {
                                const float4 x = *reinterpret_cast<const float4*>(rr + j);
                                v[j] = fmaf(g.alpha, v[j], x.x); v[j + 1] = fmaf(g.alpha, v[j + 1], x.y);
                                v[j + 2] = fmaf(g.alpha, v[j + 2], x.z); v[j + 3] = fmaf(g.alpha, v[j + 3], x.w);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < CW; ++j) if (cbase + j < nout) v[j] = fmaf(g.alpha, v[j], rr[j]);
                        }
                    } else if (g.alpha != 1.f) {
#pragma unroll
                        for (int j = 0; j < CW; ++j) v[j] *= g.alpha;
                    }
                    if constexpr (!OUTBF16) {
                        float* o = reinterpret_cast<float*>(g.out) + (int64_t)row * g.ldo + cbase;
                        if (full && (g.ldo % 4 == 0)) {
#pragma unroll
                            for (int j = 0; j < CW; j += 4)
                                *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < CW; ++j) if (cbase + j < nout) o[j] = v[j];
                        }
                    }
                    __nv_bfloat16* ob = nullptr;
                    int64_t ldb = 0;
                    if constexpr (OUTBF16) { ob = reinterpret_cast<__nv_bfloat16*>(g.out); ldb = g.ldo; }
                    else if (g.out2) { ob = reinterpret_cast<__nv_bfloat16*>(g.out2); ldb = g.ldo2; }
                    if (ob) {
                        __nv_bfloat16* o = ob + (int64_t)row * ldb + cbase;
                        if (full && (ldb % 8 == 0)) {
#pragma unroll
                            for (int j = 0; j < CW; j += 8) {
                                uint4 pk;
                                pk.x = pack_bf16(v[j], v[j + 1]); pk.y = pack_bf16(v[j + 2], v[j + 3]);
                                pk.z = pack_bf16(v[j + 4], v[j + 5]); pk.w = pack_bf16(v[j + 6], v[j + 7]);
                                *reinterpret_cast<uint4*>(o + j) = pk;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < CW; ++j) if (cbase + j < nout) o[j] = __float2bfloat16_rn(v[j]);
                        }
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty_bar[as]);
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc::tc_fence_after();
        tc::tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

template <int BN>
constexpr int stages_for() {
    // 227 KiB usable; keep ~3 KiB for barriers/alignment slack
    return ((224 * 1024) / (BM * KBYTES + BN * KBYTES)) > 8 ? 8 : ((224 * 1024) / (BM * KBYTES + BN * KBYTES));
}

template <typename TIn, int BN, int ACT, bool GLU, bool RES, bool OUTBF16>
nsp_status launch_gemm(const GemmMaps& maps, const GemmArgs& g, cudaStream_t st) {
    constexpr int STAGES = stages_for<BN>();
    constexpr size_t smem = (size_t)STAGES * (BM * KBYTES + BN * KBYTES) + 1024 /*align*/ + 256 /*barriers*/;
    auto kern = gemm_kernel<TIn, BN, STAGES, ACT, GLU, RES, OUTBF16>;
    static bool attr_set = false;
    if (!attr_set) {
        NSP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int nout = GLU ? g.N / 2 : g.N;
    const int bn_out = GLU ? BN / 2 : BN;
    const int tiles = ceil_div(g.M, BM) * ceil_div(nout, bn_out);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    launch_k(kern, dim3(grid), dim3(NTHREADS), smem, st, maps, g);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

// epilogue flavour dispatch for one operand type / tile width
template <typename TIn, int BN>
nsp_status dispatch_epi(const GemmMaps& maps, const GemmArgs& g, int glu, int act, int out_bf16, cudaStream_t st) {
    const bool res = g.residual != nullptr;
    if (glu) {
        if constexpr (BN == 128) {
            if (res) { set_error("gemm: GLU with residual is not instantiated"); return NSP_ERR_UNSUPPORTED; }
            return out_bf16 ? launch_gemm<TIn, 128, ACT_NONE, true, false, true>(maps, g, st)
                            : launch_gemm<TIn, 128, ACT_NONE, true, false, false>(maps, g, st);
        } else {
            set_error("gemm: GLU needs the 128-wide tile"); return NSP_ERR_INVALID;
        }
    }
    if (res) {
        if (act != ACT_NONE || out_bf16) { set_error("gemm: residual epilogue is fp32-out without activation"); return NSP_ERR_UNSUPPORTED; }
        return launch_gemm<TIn, BN, ACT_NONE, false, true, false>(maps, g, st);
    }
    switch (act) {
        case ACT_NONE: return out_bf16 ? launch_gemm<TIn, BN, ACT_NONE, false, false, true>(maps, g, st)
                                       : launch_gemm<TIn, BN, ACT_NONE, false, false, false>(maps, g, st);
        case ACT_RELU: return out_bf16 ? launch_gemm<TIn, BN, ACT_RELU, false, false, true>(maps, g, st)
                                       : launch_gemm<TIn, BN, ACT_RELU, false, false, false>(maps, g, st);
        case ACT_SWISH: return out_bf16 ? launch_gemm<TIn, BN, ACT_SWISH, false, false, true>(maps, g, st)
                                        : launch_gemm<TIn, BN, ACT_SWISH, false, false, false>(maps, g, st);
        case ACT_GELU: return out_bf16 ? launch_gemm<TIn, BN, ACT_GELU, false, false, true>(maps, g, st)
                                       : launch_gemm<TIn, BN, ACT_GELU, false, false, false>(maps, g, st);
        case ACT_GELU_TANH: return out_bf16 ? launch_gemm<TIn, BN, ACT_GELU_TANH, false, false, true>(maps, g, st)
                                            : launch_gemm<TIn, BN, ACT_GELU_TANH, false, false, false>(maps, g, st);
    }
    set_error("gemm: act=%d", act);
    return NSP_ERR_INVALID;
}

}  // namespace

// gemm_tma_epi.cu: same GEMM with a shared-memory staged TMA-store epilogue
nsp_status gemm_ts_dispatch(int mode, const void* a, int64_t lda, const void* w, int64_t ldw, int M, int N, int K, int glu,
                            int act, const float* bias, const float* residual, int64_t ldr, float alpha, void* out, int64_t ldo,
                            int out_bf16, void* pre, int64_t ldpre, int bn1, cudaStream_t st, bool* handled);

// Kernel-selection policy, not per-call state: 2 (TMA-store epilogue + CTA pairs where the problem fills the pairs) is what
// round 2 measured fastest on every shape class (21.6 vs 24.0 ms per training step; profiles/README.md) and is the default.
// 0 / 1 remain selectable (nsp_set_gemm_epilogue, NSP_GEMM_EPILOGUE) only so that the tests can pin each kernel family.
static int g_epilogue_mode = 2;      // 0 = per-thread vector stores, 1 = TMA-store epilogue where its envelope allows,
                                     // 2 = 1 + CTA pairs (256-row tiles) for the large problems
void set_gemm_epilogue_mode(int mode) { g_epilogue_mode = mode; }
int gemm_epilogue_mode() { return g_epilogue_mode; }

// Internal entry used by the C ABI wrappers (c_api.cu).  precision: 0 = bf16 operands, 1 = tf32 (single pass),
// 2 = "fp32" (3 tf32 segments over hi/lo splits; a/w point at the hi parts, a_lo/w_lo at the lo parts).
nsp_status gemm_dispatch(int precision, const void* a, const void* a_lo, int64_t lda, const void* w, const void* w_lo,
                         int64_t ldw, int M, int N, int K, int glu, int act, const float* bias,
                         const float* residual, int64_t ldr, float alpha, void* out, int64_t ldo, int out_bf16,
                         void* out2, int64_t ldo2, void* pre, int64_t ldpre, cudaStream_t st) {
    NSP_CHECK_ARG(a && w && out, "gemm: null pointer");
    NSP_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
    NSP_CHECK_ARG(precision >= 0 && precision <= 2, "gemm: precision=%d", precision);
    NSP_CHECK_ARG(!glu || (N % 2 == 0 && (N / 2) % 8 == 0), "gemm: GLU needs N/2 multiple of 8 (N=%d)", N);
    NSP_CHECK_ARG(precision != 2 || (a_lo && w_lo), "gemm: fp32 mode needs the lo splits");
    const bool bf16 = precision == 0;
    const int es = bf16 ? 2 : 4;
    GemmMaps maps;
    GemmArgs g;
    g.M = M; g.N = N; g.K = K; g.bias = bias; g.residual = residual; g.ldr = ldr;
    g.alpha = alpha; g.out = out; g.ldo = ldo; g.out2 = out2; g.ldo2 = ldo2;
    g.pre = pre; g.ldpre = ldpre;
    g.nseg = precision == 2 ? 3 : 1;
    NSP_CHECK_ARG(act >= 0 && act <= 4, "gemm: act=%d", act);
    NSP_CHECK_ARG(!(out2 && out_bf16), "gemm: out2 is only meaningful with an fp32 primary output");
    const int nout = glu ? N / 2 : N;
    // tile width: 128 unless the problem is too small to fill the machine, then 64 (GLU always 128 = 64 value + 64 gate)
    // 256 for wide outputs: halves the A re-reads per FLOP (the kernel is L2->SMEM bound, profiles/r01_gemm.md)
    int BN = 128;
    if (!glu && (int64_t)ceil_div(M, BM) * ceil_div(nout, 128) < num_sms() && nout > 64) BN = 64;
    else if (!glu && nout % 256 == 0 && (int64_t)ceil_div(M, BM) * (nout / 256) >= 3 * (int64_t)num_sms()) BN = 256;
    if (bf16 && g_epilogue_mode >= 1 && !out2) {
        bool handled = false;
        const nsp_status r = gemm_ts_dispatch(g_epilogue_mode, a, lda, w, ldw, M, N, K, glu, act, bias, residual, ldr, alpha, out, ldo, out_bf16,
                                              pre, ldpre, BN, st, &handled);
        if (handled) return r;
    }
    const uint32_t box_b = glu ? (uint32_t)(BN / 2) : (uint32_t)BN;
    const void* as[3] = {a, a_lo, a};
    const void* ws[3] = {w, w, w_lo};
    for (int s = 0; s < g.nseg; ++s) {
        if (!make_tmap_2d(&maps.a[s], as[s], es, bf16, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM)) return NSP_ERR_INVALID;
        if (!make_tmap_2d(&maps.b[s], ws[s], es, bf16, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, box_b)) return NSP_ERR_INVALID;
    }
    if (bf16) {
        if (BN == 256) return dispatch_epi<__nv_bfloat16, 256>(maps, g, glu, act, out_bf16, st);
        return BN == 128 ? dispatch_epi<__nv_bfloat16, 128>(maps, g, glu, act, out_bf16, st)
                         : dispatch_epi<__nv_bfloat16, 64>(maps, g, glu, act, out_bf16, st);
    }
    if (BN == 256) return dispatch_epi<float, 256>(maps, g, glu, act, out_bf16, st);
    return BN == 128 ? dispatch_epi<float, 128>(maps, g, glu, act, out_bf16, st)
                     : dispatch_epi<float, 64>(maps, g, glu, act, out_bf16, st);
}

}  // namespace nsp
