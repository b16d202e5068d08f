// Relative-position multi-head self-attention on tcgen05 (bf16 operands, fp32 softmax), d_k = 64, no XL biases,
// relative term restricted to the clamped table (clamp_len in [1, 15]) or absent -- the Conformer recipes' case.
// Other shapes run on the exact CUDA-core kernel in attention_simt.cu.
//
// Replaces  RelativeMultiheadAttentionMechanism.forward  modules/relative_multihead_attention.py:178-215
//           (AC/BD einsums, _rel_shift gather, scale, masked_fill, softmax, aw.v) -- see attention_simt.cu for the
//           semantics kept (finfo.min masking, symmetric |i-j| distance, clamp).
//
// One CTA = (utterance b, head h, 128-query tile); two CTAs are resident per SM so that one CTA's softmax (MUFU bound)
// overlaps the other's MMAs.  Per 128-key tile:
//   warp 0  TMA: K tile and V tile ([128 x 64] bf16, 128B swizzle) from the fused QKV buffer (3-D maps: column, t, b)
//   warp 1  one thread: S = Q K^T (4 x tcgen05.mma M128 N128 K16) -> TMEM[0,128);  O_j = P V (8 x M128 N64 K16,
//           V used as an MN-major B operand straight from its row-major tile) -> TMEM[128,192)
//   warps 2-5  thread = query row: tcgen05.ld S in 32-column chunks, add the relative term
//           (BD_raw = Q R^T computed once per CTA by 4 MMAs into TMEM[192,208), gathered by min(|i-j|, clamp)),
//           mask from device-side lengths, online softmax in the exp2 domain, P (bf16) written to shared memory in the
//           K-major 128B-swizzled layout the PV MMA reads; O accumulates in registers with the usual rescale.
#include <float.h>
#include "tc_common.cuh"

namespace nsp {
namespace {

constexpr int QT = 128, KT = 128, DK = 64;
constexpr int TILE_BYTES = 128 * 128;                 // [128 rows x 64 bf16]
constexpr int NSM_WARPS = 8;                        // softmax warps: two per TMEM lane quadrant (key halves)
constexpr int NTHREADS = 64 + 32 * NSM_WARPS;

struct AttnTcArgs {
    const int32_t* klens;
    __nv_bfloat16* out; int64_t ldo;
    int B, H, Tq, Tk;
    int has_rel, clamp_len;
    int causal, lookahead, chunk_c, chunk_l;
    float scale_log2;                                 // log2(e) / sqrt(dk)
    float* stats;                                     // optional [B,H,Tq,2]: (row max in the log2 domain, 1 / row sum) for the backward
};

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void tmem_ld_32x32_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

// MN-major B operand (V tile: rows = keys (K), 128 B of dk (N) per row), 128B swizzle: 8-key groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc_sw128_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                           // LBO (MN-group stride): unused for N = 64
    d |= (uint64_t)(1024 >> 4) << 32;                 // SBO: stride between 8-row K groups
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(NTHREADS, 2) attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q,
                                                              const __grid_constant__ CUtensorMap tmap_k,
                                                              const __grid_constant__ CUtensorMap tmap_v,
                                                              const __grid_constant__ CUtensorMap tmap_r,
                                                              const AttnTcArgs a) {
    pdl_launch_dependents();      // PDL: the next kernel may start its prologue; ours overlaps the previous kernel's tail
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                               // 16 KiB
    uint8_t* sK = sQ + TILE_BYTES;                    // 16 KiB
    uint8_t* sV = sK + TILE_BYTES;                    // 16 KiB
    uint8_t* sP = sV + TILE_BYTES;                    // 2 x 16 KiB (keys 0-63, 64-127)
    uint8_t* sR = sP + 2 * TILE_BYTES;                // 16 rows x 128 B = 2 KiB
    float* sBD = reinterpret_cast<float*>(sR + 2048); // [128][17]
    float* sMX = sBD + 128 * 17;                      // [2][128] partial row maxima of the two key halves
    float* sL = sMX + 256;                            // [2][128] partial row sums (final combine)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sL + 256);
    uint64_t* q_full = bars + 0; uint64_t* k_full = bars + 1; uint64_t* k_empty = bars + 2;
    uint64_t* v_full = bars + 3; uint64_t* v_empty = bars + 4; uint64_t* s_full = bars + 5;
    uint64_t* p_full = bars + 6; uint64_t* o_full = bars + 7; uint64_t* bd_full = bars + 8;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qtiles = (a.Tq + QT - 1) / QT;
    const int qt = blockIdx.x % qtiles;
    const int h = (blockIdx.x / qtiles) % a.H;
    const int b = blockIdx.x / (qtiles * a.H);
    const int i0 = qt * QT;
    const int mlen = a.Tk - a.Tq;
    const int klen = min(max(a.klens[b], 0), a.Tk);
    // keys that can be visible to this query tile: [0, kmax)
    int kmax = klen;
    if (a.causal) kmax = min(kmax, mlen + i0 + QT - 1 + a.lookahead + 1);
    kmax = max(kmax, 1);
    const int ntiles = (kmax + KT - 1) / KT;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap_q); tc::tma_prefetch_desc(&tmap_k); tc::tma_prefetch_desc(&tmap_v);
        if (a.has_rel) tc::tma_prefetch_desc(&tmap_r);
    }
    if (warp == 1 && lane == 0) {
        tc::mbar_init(q_full, 1); tc::mbar_init(k_full, 1); tc::mbar_init(k_empty, 1);
        tc::mbar_init(v_full, 1); tc::mbar_init(v_empty, 1); tc::mbar_init(s_full, 1);
        tc::mbar_init(p_full, 32 * NSM_WARPS); tc::mbar_init(o_full, 1); tc::mbar_init(bd_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<256>(tmem_holder);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();                    // the previous grid is complete: operands / residuals / outputs may be touched from here
    const uint32_t tm_S = tmem_base, tm_O = tmem_base + 128, tm_BD = tmem_base + 192;

    if (warp == 0) {
        if (tc::elect_one()) {
            tc::mbar_arrive_expect_tx(q_full, TILE_BYTES + (a.has_rel ? 2048 : 0));
            tc::tma_load_3d(sQ, &tmap_q, q_full, h * DK, i0, b);
            if (a.has_rel) tc::tma_load_2d(sR, &tmap_r, q_full, h * DK, 0);
            for (int j = 0; j < ntiles; ++j) {
                const uint32_t ph = j & 1;
                tc::mbar_wait(k_empty, ph ^ 1);
                tc::mbar_arrive_expect_tx(k_full, TILE_BYTES);
                tc::tma_load_3d(sK, &tmap_k, k_full, h * DK, j * KT, b);
                tc::mbar_wait(v_empty, ph ^ 1);
                tc::mbar_arrive_expect_tx(v_full, TILE_BYTES);
                tc::tma_load_3d(sV, &tmap_v, v_full, h * DK, j * KT, b);
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint32_t idesc_s = tc::make_idesc(1u, 128, 128);
            constexpr uint32_t idesc_bd = tc::make_idesc(1u, 128, 16);
            constexpr uint32_t idesc_pv = tc::make_idesc(1u, 128, 64) | (1u << 16);      // B operand MN-major
            tc::mbar_wait(q_full, 0);
            tc::tc_fence_after();
            const uint64_t qdesc = tc::make_smem_desc_sw128(tc::smem_u32(sQ));
            if (a.has_rel) {
                const uint64_t rdesc = tc::make_smem_desc_sw128(tc::smem_u32(sR));
#pragma unroll
                for (int k = 0; k < 4; ++k) tc::umma_f16(tm_BD, qdesc + 2 * k, rdesc + 2 * k, idesc_bd, k > 0);
                tc::umma_commit(bd_full);
            }
            const uint64_t kdesc = tc::make_smem_desc_sw128(tc::smem_u32(sK));
            const uint64_t vdesc = make_smem_desc_sw128_mn(tc::smem_u32(sV));
            const uint64_t pdesc0 = tc::make_smem_desc_sw128(tc::smem_u32(sP));
            const uint64_t pdesc1 = tc::make_smem_desc_sw128(tc::smem_u32(sP + TILE_BYTES));
            for (int j = 0; j < ntiles; ++j) {
                const uint32_t ph = j & 1;
                tc::mbar_wait(k_full, ph);
                tc::tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k) tc::umma_f16(tm_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0);
                tc::umma_commit(s_full);
                tc::umma_commit(k_empty);
                tc::mbar_wait(p_full, ph);
                tc::mbar_wait(v_full, ph);
                tc::tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    // P: K-major, 16 keys = 32 B inside the 128 B swizzle row (second half tile after 4 steps);
                    // V: MN-major, 16 keys = 16 rows of 128 B = 2048 B
                    const uint64_t pd = (k < 4 ? pdesc0 : pdesc1) + 2 * (k & 3);
                    tc::umma_f16(tm_O, pd, vdesc + (uint64_t)((k * 2048) >> 4), idesc_pv, k > 0);
                }
                tc::umma_commit(o_full);
                tc::umma_commit(v_empty);
            }
        }
    } else {
        // two threads per query row: `half` 0/1 owns keys [0,64) / [64,128) of every key tile and d_k columns
        // [0,32) / [32,64) of the output; they meet through shared memory + a named barrier over the 8 softmax warps
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = q * 32 + lane;                       // query row inside the tile == TMEM lane
        const int i = i0 + row;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        auto sm_barrier = [&]() { asm volatile("bar.sync 1, %0;" :: "n"(32 * NSM_WARPS) : "memory"); };
        float bd_far = 0.f;
        const int clamp = a.clamp_len;
        if (a.has_rel) {
            tc::mbar_wait(bd_full, 0);
            tc::tc_fence_after();
            if (half == 0) {
                uint32_t r16[16];
                tmem_ld_32x32_x16(tm_BD + lane_addr, r16);
                tc::tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 16; ++c) sBD[row * 17 + c] = __uint_as_float(r16[c]);
            }
            sm_barrier();
            bd_far = sBD[row * 17 + clamp];                  // every |i-j| >= clamp_len shares this value
        }
        constexpr int DH = DK / 2;
        float O[DH];
#pragma unroll
        for (int c = 0; c < DH; ++c) O[c] = 0.f;
        float m_run = -FLT_MAX, l_run = 0.f;                 // l_run: this half's share of the row sum
        const int iw0 = i0 + q * 32, iw1 = iw0 + 31;
        const float bdfs = bd_far * a.scale_log2;

        for (int j = 0; j < ntiles; ++j) {
            const uint32_t ph = j & 1;
            const int j0 = j * KT + half * (KT / 2);         // first key of this thread's half tile
            const uint32_t s_col = (uint32_t)(half * (KT / 2));
            tc::mbar_wait(s_full, ph);
            tc::tc_fence_after();
            auto score = [&](float raw, int jj) -> float {
                float s = raw;
                if (a.has_rel) {
                    int d = mlen + i - jj; d = d < 0 ? -d : d;
                    s += (d >= clamp) ? bd_far : sBD[row * 17 + d];
                }
                s *= a.scale_log2;
                bool vis = jj < klen;
                if (a.causal) vis = vis && (jj <= mlen + i + a.lookahead);
                if (a.chunk_c > 0) {
                    int cs = ((mlen + i) / a.chunk_c) * a.chunk_c;
                    vis = vis && (jj >= cs - a.chunk_l) && (jj < cs + a.chunk_c);
                }
                if (!vis) s = -FLT_MAX;
                if (jj >= a.Tk) s = -INFINITY;
                return s;
            };
            // warp-uniform fast path per 32-key chunk: every key visible to every row of this warp and the chunk
            // outside the clamped band -> score = raw * scale + const (one FFMA), no per-element mask / gather logic
            auto chunk_fast = [&](int c) -> bool {
                const int jc0 = j0 + c, jc1 = jc0 + 31;
                const bool vis = (jc1 < klen) && !a.causal && (a.chunk_c == 0);
                const bool far = !a.has_rel || (jc1 <= mlen + iw0 - clamp) || (jc0 >= mlen + iw1 + clamp);
                return vis && far;
            };
            // ---- pass 1: row max over this half's 64 keys ----
            float mx = -FLT_MAX;
#pragma unroll 1
            for (int c = 0; c < KT / 2; c += 32) {
                uint32_t r[32];
                tc::tmem_ld_32x32(tm_S + lane_addr + s_col + (uint32_t)c, r);
                tc::tmem_ld_wait();
                if (chunk_fast(c)) {
                    float mr = -FLT_MAX;
#pragma unroll
                    for (int e = 0; e < 32; ++e) mr = fmaxf(mr, __uint_as_float(r[e]));
                    mx = fmaxf(mx, fmaf(mr, a.scale_log2, bdfs));
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) mx = fmaxf(mx, score(__uint_as_float(r[e]), j0 + c + e));
                }
            }
            sMX[half * 128 + row] = mx;
            sm_barrier();
            const float m_new = fmaxf(m_run, fmaxf(mx, sMX[(half ^ 1) * 128 + row]));
            const float corr = ex2(m_run - m_new);
            float rsum = 0.f;
            // ---- pass 2: probabilities -> bf16 P half tile (K-major, 128B swizzle) ----
#pragma unroll 1
            for (int c = 0; c < KT / 2; c += 32) {
                uint32_t r[32];
                tc::tmem_ld_32x32(tm_S + lane_addr + s_col + (uint32_t)c, r);
                tc::tmem_ld_wait();
                uint32_t pk[16];
                if (chunk_fast(c)) {
                    const float off = bdfs - m_new;
#pragma unroll
                    for (int e = 0; e < 32; e += 2) {
                        float p0 = ex2(fmaf(__uint_as_float(r[e]), a.scale_log2, off));
                        float p1 = ex2(fmaf(__uint_as_float(r[e + 1]), a.scale_log2, off));
                        __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
                        rsum += __bfloat162float(pb.x) + __bfloat162float(pb.y);
                        pk[e >> 1] = *reinterpret_cast<uint32_t*>(&pb);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 32; e += 2) {
                        float s0 = score(__uint_as_float(r[e]), j0 + c + e);
                        float s1 = score(__uint_as_float(r[e + 1]), j0 + c + e + 1);
                        float p0 = (s0 == -INFINITY) ? 0.f : ex2(s0 - m_new);
                        float p1 = (s1 == -INFINITY) ? 0.f : ex2(s1 - m_new);
                        __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
                        // the row sum is accumulated from the ROUNDED probabilities so that P V and l stay consistent
                        rsum += __bfloat162float(pb.x) + __bfloat162float(pb.y);
                        pk[e >> 1] = *reinterpret_cast<uint32_t*>(&pb);
                    }
                }
                uint8_t* hrow = sP + half * TILE_BYTES + row * 128;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int unit = ((c & 32) >> 3) + u;                 // 16-byte unit inside the 128 B row
                    uint4 v4 = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                    *reinterpret_cast<uint4*>(hrow + ((unit ^ (row & 7)) << 4)) = v4;
                }
            }
            l_run = l_run * corr + rsum;
            m_run = m_new;
            tc::tc_fence_before();
            tc::fence_proxy_async_smem();                                 // generic-proxy smem writes -> visible to the MMA
            tc::mbar_arrive(p_full);
            // ---- O += P V (this thread keeps d_k columns [half*32, half*32+32)) ----
            tc::mbar_wait(o_full, ph);
            tc::tc_fence_after();
            {
                uint32_t r[32];
                tc::tmem_ld_32x32(tm_O + lane_addr + (uint32_t)(half * DH), r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < DH; ++e) O[e] = O[e] * corr + __uint_as_float(r[e]);
            }
            tc::tc_fence_before();
        }
        sL[half * 128 + row] = l_run;
        sm_barrier();
        if (i < a.Tq) {
            const float inv = 1.f / (l_run + sL[(half ^ 1) * 128 + row]);
            if (a.stats && half == 0) {
                float* st = a.stats + (((int64_t)b * a.H + h) * a.Tq + i) * 2;
                st[0] = m_run; st[1] = inv;
            }
            __nv_bfloat16* o = a.out + ((int64_t)b * a.Tq + i) * a.ldo + (int64_t)h * DK + half * DH;
#pragma unroll
            for (int c = 0; c < DH; c += 8) {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(O[c] * inv, O[c + 1] * inv);
                __nv_bfloat162 p1 = __floats2bfloat162_rn(O[c + 2] * inv, O[c + 3] * inv);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(O[c + 4] * inv, O[c + 5] * inv);
                __nv_bfloat162 p3 = __floats2bfloat162_rn(O[c + 6] * inv, O[c + 7] * inv);
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
                pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
                *reinterpret_cast<uint4*>(o + c) = pk;
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc<256>(tmem_base); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_map3(EncodeTiledFn enc, CUtensorMap* m, const void* base, int64_t ld, int T, int B, int cols, const char* what) {
    if (((uintptr_t)base % 16) != 0 || (ld * 2) % 16 != 0) {
        set_error("attention_tc: %s must be 16-byte aligned with a 16-byte-multiple pitch", what);
        return false;
    }
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("attention_tc: tensor map (%s) failed: %d", what, (int)r); return false; }
    return true;
}

}  // namespace

bool get_tma_encode(void** fn);

// returns NSP_ERR_UNSUPPORTED (without setting a launch) when the shape is outside this kernel's envelope
nsp_status attention_tc_dispatch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                 const void* r, int64_t ldr, int rlen, const int32_t* klens, void* out, int64_t ldo,
                                 int B, int H, int Tq, int Tk, int dk, int clamp_len, int causal, int lookahead,
                                 int chunk_c, int chunk_l, float* stats, cudaStream_t st) {
    if (dk != 64) return NSP_ERR_UNSUPPORTED;
    if (r && !(clamp_len >= 1 && clamp_len <= 15)) return NSP_ERR_UNSUPPORTED;
    if (ldo % 8 != 0 || ((uintptr_t)out % 16) != 0) return NSP_ERR_UNSUPPORTED;
    void* fnp = nullptr;
    if (!get_tma_encode(&fnp)) return NSP_ERR_CUDA;
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    CUtensorMap mq, mk, mv, mr;
    const int cols = H * dk;
    if (!make_map3(enc, &mq, q, ldq, Tq, B, cols, "q")) return NSP_ERR_INVALID;
    if (!make_map3(enc, &mk, k, ldk, Tk, B, cols, "k")) return NSP_ERR_INVALID;
    if (!make_map3(enc, &mv, v, ldv, Tk, B, cols, "v")) return NSP_ERR_INVALID;
    mr = mq;
    if (r) {
        if (((uintptr_t)r % 16) != 0 || (ldr * 2) % 16 != 0) { set_error("attention_tc: r alignment"); return NSP_ERR_INVALID; }
        cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rlen};
        cuuint64_t gstr[1] = {(cuuint64_t)ldr * 2};
        cuuint32_t box[2] = {64, 16};
        cuuint32_t es[2] = {1, 1};
        CUresult rc = enc(&mr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(r), gdim, gstr, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) { set_error("attention_tc: tensor map (r) failed: %d", (int)rc); return NSP_ERR_CUDA; }
    }
    AttnTcArgs a;
    a.klens = klens; a.out = (__nv_bfloat16*)out; a.ldo = ldo; a.B = B; a.H = H; a.Tq = Tq; a.Tk = Tk;
    a.has_rel = r ? 1 : 0; a.clamp_len = r ? (clamp_len < rlen - 1 ? clamp_len : rlen - 1) : 0;
    a.causal = causal; a.lookahead = lookahead; a.chunk_c = chunk_c; a.chunk_l = chunk_l;
    a.scale_log2 = 1.4426950408889634f / sqrtf((float)dk);
    a.stats = stats;
    const size_t smem = 1024 + 5 * TILE_BYTES + 2048 + (128 * 17 + 512) * sizeof(float) + 128;
    static bool attr = false;
    if (!attr) { NSP_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    const int qtiles = ceil_div(Tq, QT);
    launch_k(attn_tc_kernel, dim3((unsigned)(B * H * qtiles)), dim3(NTHREADS), smem, st, mq, mk, mv, mr, a);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

}  // namespace nsp
