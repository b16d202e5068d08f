// Backward of the (relative-position) self-attention core on tcgen05 (bf16 operands, fp32 statistics), d_k = 64, no XL
// biases, relative term clamped (clamp_len in [1,15]) or absent, Tq == Tk -- the training counterpart of attention_tc.cu.
// Other shapes use the CUDA-core kernels in attention_bwd.cu.
//
// The reference obtains this from autograd over relative_multihead_attention.py:178-215 (see attention_bwd.cu for the
// formulas).  One CTA = (utterance b, head h, 128-key tile); it walks the 128-query tiles and per tile pair runs five
// tensor-core GEMMs, with the roles transposed with respect to the forward kernel so that dK / dV accumulate in TMEM:
//   S^T  = K Q^T        [keys x queries]   (TMEM, lane = key)        dP^T = V dO^T      [keys x queries]
//   P^T  = exp2(S^T - m_i) / l_i , dS^T = P^T (dP^T - D_i) / sqrt(dk)   (thread = key row, statistics per query column
//          come from the forward kernel; written as bf16 K-major tiles into shared memory)
//   dV  += P^T dO       dK  += dS^T Q     (A = the bf16 tiles, B = dO / Q tiles read MN-major as they sit in smem)
//   dQ_t = dS K         (A = the SAME dS^T tile read as an MN-major operand, B = K tile MN-major) -> red.global.add
// Relative term: BD = Q R^T (one N=16 MMA per query tile) is gathered by min(|i-j|, clamp) when the scores are rebuilt;
// its gradient needs W[i][d] = sum_{j: dist(i,j)=d} dS_ij only for d < clamp (the softmax Jacobian makes every row of
// dS sum to zero, so W[i][clamp] = -sum_{d<clamp} W[i][d]); those few near-diagonal terms go to global memory by
// atomicAdd and a small finishing kernel forms dq = dQ + W R (bf16) and dR += W^T q.
#include <float.h>
#include "tc_common.cuh"

namespace nsp {
namespace {

constexpr int QT = 128, KT = 128, DK = 64;
constexpr int TILE = 128 * 128;                       // bytes of a [128 x 64] bf16 tile
constexpr int NCW = 16;                               // compute warps (four per TMEM lane quadrant: 32 query columns each)
constexpr int NTHREADS = 64 + 32 * NCW;

struct BwdTcArgs {
    const int32_t* klens;
    const float* stats;        // [B,H,T,2] (m in the log2 domain, 1/l) from the forward kernel
    const float* dsum;         // [B,H,T]   D_i = dO_i . O_i
    float* dq_acc;             // [B*T, H*64] fp32, zero-initialised
    float* w;                  // [B,H,T,16] fp32, zero-initialised (near-band sums of dS)
    __nv_bfloat16* dk; int64_t lddk;
    __nv_bfloat16* dv; int64_t lddv;
    int B, H, T;
    int has_rel, clamp;
    int causal, lookahead, chunk_c, chunk_l;
    float scale_log2, inv_scale;
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
// MN-major operand tile made of 128-byte rows (rows = the MMA's K index), 128B swizzle.
// lbo = byte stride between 64-element chunks of the M/N index (unused when that extent is 64), SBO = 8-row groups.
__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Bounded mbarrier wait (tc_common.cuh: bare spin + one counter, the clock is read once per 4096 failed polls).
__device__ __forceinline__ void bwait(uint64_t* bar, uint32_t parity, int) { tc::mbar_wait(bar, parity); }

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q,
                                                                  const __grid_constant__ CUtensorMap tmap_k,
                                                                  const __grid_constant__ CUtensorMap tmap_v,
                                                                  const __grid_constant__ CUtensorMap tmap_do,
                                                                  const __grid_constant__ CUtensorMap tmap_r,
                                                                  const BwdTcArgs a) {
    pdl_launch_dependents();      // PDL: the next kernel may start its prologue; ours overlaps the previous kernel's tail
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sK = smem;                               // 16 KiB
    uint8_t* sV = sK + TILE;                          // 16 KiB
    uint8_t* sQ = sV + TILE;                          // 2 x 16 KiB
    uint8_t* sDO = sQ + 2 * TILE;                     // 2 x 16 KiB
    uint8_t* sPT = sDO + 2 * TILE;                    // 2 x 16 KiB: P^T, queries [0,64) | [64,128)
    uint8_t* sDS = sPT + 2 * TILE;                    // 2 x 16 KiB: dS^T
    uint8_t* sR = sDS + 2 * TILE;                     // 16 rows x 128 B
    float* sBD = reinterpret_cast<float*>(sR + 2048); // [128 queries][17]
    float* sST = sBD + 128 * 17;                      // [3][128]: m, 1/l, D of the current query tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(sST + 3 * 128);
    uint64_t* kv_full = bars + 0;
    uint64_t* qd_full = bars + 1;                     // [2]
    uint64_t* qd_empty = bars + 3;                    // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* p_full = bars + 6;
    uint64_t* o_full = bars + 7;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ktiles = (a.T + KT - 1) / KT;
    const int kt = blockIdx.x % ktiles;
    const int h = (blockIdx.x / ktiles) % a.H;
    const int b = blockIdx.x / (ktiles * a.H);
    const int j0 = kt * KT;
    const int klen = min(max(a.klens[b], 0), a.T);
    const int qtiles = (a.T + QT - 1) / QT;
    // causal: query tiles entirely before this key tile see none of its keys
    int t_first = 0;
    if (a.causal) t_first = max(0, (j0 - a.lookahead) / QT);
    const int ntiles = qtiles - t_first;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap_q); tc::tma_prefetch_desc(&tmap_k); tc::tma_prefetch_desc(&tmap_v);
        tc::tma_prefetch_desc(&tmap_do);
        if (a.has_rel) tc::tma_prefetch_desc(&tmap_r);
    }
    if (warp == 1 && lane == 0) {
        tc::mbar_init(kv_full, 1);
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&qd_full[i], 1); tc::mbar_init(&qd_empty[i], 1); }
        tc::mbar_init(s_full, 1); tc::mbar_init(p_full, 32 * NCW); tc::mbar_init(o_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<512>(tmem_holder);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();                    // the previous grid is complete: operands / residuals / outputs may be touched from here
    const uint32_t tm_ST = tmem_base, tm_DP = tmem_base + 128, tm_DV = tmem_base + 256, tm_DK = tmem_base + 320;
    const uint32_t tm_DQ = tmem_base + 384, tm_BD = tmem_base + 448;

    if (warp == 0) {
        if (tc::elect_one()) {
            tc::mbar_arrive_expect_tx(kv_full, 2 * TILE + (a.has_rel ? 2048 : 0));
            tc::tma_load_3d(sK, &tmap_k, kv_full, h * DK, j0, b);
            tc::tma_load_3d(sV, &tmap_v, kv_full, h * DK, j0, b);
            if (a.has_rel) tc::tma_load_2d(sR, &tmap_r, kv_full, h * DK, 0);
            for (int t = 0; t < ntiles; ++t) {
                const int buf = t & 1;
                bwait(&qd_empty[buf], ((t >> 1) & 1) ^ 1, 1);
                tc::mbar_arrive_expect_tx(&qd_full[buf], 2 * TILE);
                tc::tma_load_3d(sQ + buf * TILE, &tmap_q, &qd_full[buf], h * DK, (t_first + t) * QT, b);
                tc::tma_load_3d(sDO + buf * TILE, &tmap_do, &qd_full[buf], h * DK, (t_first + t) * QT, b);
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint32_t idesc_s = tc::make_idesc(1u, 128, 128);                              // K-major x K-major
            constexpr uint32_t idesc_bd = tc::make_idesc(1u, 128, 16);
            constexpr uint32_t idesc_acc = tc::make_idesc(1u, 128, 64) | (1u << 16);                // B MN-major
            constexpr uint32_t idesc_dq = tc::make_idesc(1u, 128, 64) | (1u << 15) | (1u << 16);    // A and B MN-major
            bwait(kv_full, 0, 2);
            tc::tc_fence_after();
            const uint64_t kdesc = tc::make_smem_desc_sw128(tc::smem_u32(sK));
            const uint64_t vdesc = tc::make_smem_desc_sw128(tc::smem_u32(sV));
            const uint64_t kdesc_mn = desc_mn(tc::smem_u32(sK), 16);
            const uint64_t rdesc = tc::make_smem_desc_sw128(tc::smem_u32(sR));
            const uint64_t pt0 = tc::make_smem_desc_sw128(tc::smem_u32(sPT)), pt1 = tc::make_smem_desc_sw128(tc::smem_u32(sPT + TILE));
            const uint64_t ds0 = tc::make_smem_desc_sw128(tc::smem_u32(sDS)), ds1 = tc::make_smem_desc_sw128(tc::smem_u32(sDS + TILE));
            const uint64_t ds_mn = desc_mn(tc::smem_u32(sDS), TILE);       // queries [0,64) | [64,128) are TILE bytes apart
            for (int t = 0; t < ntiles; ++t) {
                const int buf = t & 1;
                const uint32_t ph = t & 1;
                bwait(&qd_full[buf], (t >> 1) & 1, 3);
                tc::tc_fence_after();
                const uint64_t qdesc = tc::make_smem_desc_sw128(tc::smem_u32(sQ + buf * TILE));
                const uint64_t dodesc = tc::make_smem_desc_sw128(tc::smem_u32(sDO + buf * TILE));
                const uint64_t q_mn = desc_mn(tc::smem_u32(sQ + buf * TILE), 16);
                const uint64_t do_mn = desc_mn(tc::smem_u32(sDO + buf * TILE), 16);
                if (a.has_rel) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc::umma_f16(tm_BD, qdesc + 2 * k, rdesc + 2 * k, idesc_bd, k > 0);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) tc::umma_f16(tm_ST, kdesc + 2 * k, qdesc + 2 * k, idesc_s, k > 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) tc::umma_f16(tm_DP, vdesc + 2 * k, dodesc + 2 * k, idesc_s, k > 0);
                tc::umma_commit(s_full);
                bwait(p_full, ph, 4);
                tc::tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    // A (K-major, K = queries): 16 queries = 32 B inside a 128 B row, second half tile after 4 steps;
                    // B (MN-major, K = queries): 16 query rows of 128 B = 2048 B
                    const uint64_t koffb = (uint64_t)((k * 2048) >> 4);
                    tc::umma_f16(tm_DV, (k < 4 ? pt0 : pt1) + 2 * (k & 3), do_mn + koffb, idesc_acc, (t > 0 || k > 0) ? 1u : 0u);
                    tc::umma_f16(tm_DK, (k < 4 ? ds0 : ds1) + 2 * (k & 3), q_mn + koffb, idesc_acc, (t > 0 || k > 0) ? 1u : 0u);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    // dQ_t[queries x 64] = dS K: A = dS^T tile read MN-major (K = keys: 16 key rows = 2048 B per step)
                    const uint64_t koffb = (uint64_t)((k * 2048) >> 4);
                    tc::umma_f16(tm_DQ, ds_mn + koffb, kdesc_mn + koffb, idesc_dq, k > 0);
                }
                tc::umma_commit(o_full);
                tc::umma_commit(&qd_empty[buf]);
            }
        }
    } else {
        const int q = warp & 3;
        const int quarter = (warp - 2) >> 2;                 // which 32 of the 128 query columns (16 of the 64 dq / dk / dv columns)
        const int row = q * 32 + lane;                       // key row of S^T / dP^T, query row of BD / dQ_t
        const int j = j0 + row;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const int ctid = threadIdx.x - 64;                   // 0..511
        auto cbar = [&]() { asm volatile("bar.sync 1, %0;" :: "n"(32 * NCW) : "memory"); };
        const int clamp = a.clamp;
        const int64_t bh = (int64_t)b * a.H + h;
        const bool jin = j < a.T;
        const bool jvis_len = j < klen;

        for (int t = 0; t < ntiles; ++t) {
            const uint32_t ph = t & 1;
            const int i0 = (t_first + t) * QT;
            cbar();                                           // everyone is done with sST / sBD of the previous tile
            for (int e = ctid; e < 3 * 128; e += 32 * NCW) {
                const int which = e >> 7, qi = e & 127, i = i0 + qi;
                float v = 0.f;
                if (i < a.T) v = (which < 2) ? a.stats[(bh * a.T + i) * 2 + which] : a.dsum[bh * a.T + i];
                sST[e] = v;
            }
            bwait(s_full, ph, 5);
            tc::tc_fence_after();
            if (a.has_rel && quarter == 0) {
                uint32_t r16[16];
                tmem_ld16(tm_BD + lane_addr, r16);
                tc::tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 16; ++c) sBD[row * 17 + c] = __uint_as_float(r16[c]);
            }
            cbar();
            // ---- P^T and dS^T of this thread's key row for its 32 query columns (two 16-column halves) ----
#pragma unroll 1
            for (int c = 0; c < 32; c += 16) {
                uint32_t rs[16], rp[16];
                tmem_ld16(tm_ST + lane_addr + (uint32_t)(quarter * 32 + c), rs);
                tmem_ld16(tm_DP + lane_addr + (uint32_t)(quarter * 32 + c), rp);
                tc::tmem_ld_wait();
                uint32_t pk[8], dk_[8];
                const int qc0 = quarter * 32 + c;             // first query column (tile-local) of this half chunk
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    float pv[2], dv[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int qi = qc0 + e + u, i = i0 + qi;
                        float s = __uint_as_float(rs[e + u]);
                        int dist = i - j; dist = dist < 0 ? -dist : dist;
                        if (a.has_rel) s += sBD[qi * 17 + (dist < clamp ? dist : clamp)];
                        s *= a.scale_log2;
                        bool vis = jvis_len;
                        if (a.causal) vis = vis && (j <= i + a.lookahead);
                        if (a.chunk_c > 0) {
                            const int cs = (i / a.chunk_c) * a.chunk_c;
                            vis = vis && (j >= cs - a.chunk_l) && (j < cs + a.chunk_c);
                        }
                        if (!vis) s = -FLT_MAX;
                        const float p = jin ? ex2f(s - sST[qi]) * sST[128 + qi] : 0.f;
                        const float ds = (vis && jin) ? p * (__uint_as_float(rp[e + u]) - sST[256 + qi]) * a.inv_scale : 0.f;
                        if (a.has_rel && dist < clamp && ds != 0.f) atomicAdd(a.w + (bh * a.T + i) * 16 + dist, ds);
                        pv[u] = p; dv[u] = ds;
                    }
                    __nv_bfloat162 pb = __floats2bfloat162_rn(pv[0], pv[1]);
                    __nv_bfloat162 db = __floats2bfloat162_rn(dv[0], dv[1]);
                    pk[e >> 1] = *reinterpret_cast<uint32_t*>(&pb);
                    dk_[e >> 1] = *reinterpret_cast<uint32_t*>(&db);
                }
                // 128-byte rows hold 64 queries: half tile = quarter >> 1, 16-byte unit = (quarter & 1) * 4 + c / 8 + u
                uint8_t* prow = sPT + (quarter >> 1) * TILE + row * 128;
                uint8_t* drow = sDS + (quarter >> 1) * TILE + row * 128;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int unit = (quarter & 1) * 4 + (c >> 3) + u;
                    const int off = (unit ^ (row & 7)) << 4;
                    *reinterpret_cast<uint4*>(prow + off) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                    *reinterpret_cast<uint4*>(drow + off) = make_uint4(dk_[4 * u], dk_[4 * u + 1], dk_[4 * u + 2], dk_[4 * u + 3]);
                }
            }
            tc::tc_fence_before();
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(p_full);
            // ---- dQ_t: lane = query row; this thread adds 32 of the 64 columns ----
            bwait(o_full, ph, 6);
            tc::tc_fence_after();
            {
                uint32_t r[16];
                tmem_ld16(tm_DQ + lane_addr + (uint32_t)(quarter * 16), r);
                tc::tmem_ld_wait();
                const int i = i0 + row;
                if (i < a.T) {
                    float* dst = a.dq_acc + ((int64_t)b * a.T + i) * ((int64_t)a.H * DK) + h * DK + quarter * 16;
#pragma unroll
                    for (int e = 0; e < 16; e += 4)
                        red_add_v4(dst + e, __uint_as_float(r[e]), __uint_as_float(r[e + 1]), __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
                }
            }
            tc::tc_fence_before();
        }
        // ---- dK, dV of this key tile (complete after the last o_full) ----
        // tcgen05.ld is warp-collective (.sync.aligned): issue it from the whole warp, guard only the stores by `jin`
        uint32_t rv[16], rk[16];
        if (ntiles > 0) {
            tmem_ld16(tm_DV + lane_addr + (uint32_t)(quarter * 16), rv);
            tmem_ld16(tm_DK + lane_addr + (uint32_t)(quarter * 16), rk);
            tc::tmem_ld_wait();
        }
        if (jin) {
            __nv_bfloat16* ov = a.dv + ((int64_t)b * a.T + j) * a.lddv + h * DK + quarter * 16;
            __nv_bfloat16* ok = a.dk + ((int64_t)b * a.T + j) * a.lddk + h * DK + quarter * 16;
#pragma unroll
            for (int c = 0; c < 16; c += 8) {
                uint4 pv = make_uint4(0, 0, 0, 0), pkk = make_uint4(0, 0, 0, 0);     // no query tile sees this key tile: zeros
                if (ntiles > 0) {
                    __nv_bfloat162 t0, t1, t2, t3;
                    t0 = __floats2bfloat162_rn(__uint_as_float(rv[c]), __uint_as_float(rv[c + 1]));
                    t1 = __floats2bfloat162_rn(__uint_as_float(rv[c + 2]), __uint_as_float(rv[c + 3]));
                    t2 = __floats2bfloat162_rn(__uint_as_float(rv[c + 4]), __uint_as_float(rv[c + 5]));
                    t3 = __floats2bfloat162_rn(__uint_as_float(rv[c + 6]), __uint_as_float(rv[c + 7]));
                    pv.x = *reinterpret_cast<uint32_t*>(&t0); pv.y = *reinterpret_cast<uint32_t*>(&t1);
                    pv.z = *reinterpret_cast<uint32_t*>(&t2); pv.w = *reinterpret_cast<uint32_t*>(&t3);
                    t0 = __floats2bfloat162_rn(__uint_as_float(rk[c]), __uint_as_float(rk[c + 1]));
                    t1 = __floats2bfloat162_rn(__uint_as_float(rk[c + 2]), __uint_as_float(rk[c + 3]));
                    t2 = __floats2bfloat162_rn(__uint_as_float(rk[c + 4]), __uint_as_float(rk[c + 5]));
                    t3 = __floats2bfloat162_rn(__uint_as_float(rk[c + 6]), __uint_as_float(rk[c + 7]));
                    pkk.x = *reinterpret_cast<uint32_t*>(&t0); pkk.y = *reinterpret_cast<uint32_t*>(&t1);
                    pkk.z = *reinterpret_cast<uint32_t*>(&t2); pkk.w = *reinterpret_cast<uint32_t*>(&t3);
                }
                *reinterpret_cast<uint4*>(ov + c) = pv;
                *reinterpret_cast<uint4*>(ok + c) = pkk;
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc<512>(tmem_base); }
}

// D[b,h,i] = sum_c dO[b,i,h,c] * O[b,i,h,c]
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, int64_t ldo,
                                                            const __nv_bfloat16* __restrict__ dout, int64_t lddo,
                                                            float* __restrict__ dsum, int B, int H, int T) {
    pdl_entry();
    const int64_t n = (int64_t)B * T * H;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
        const int h = (int)(e % H);
        const int64_t bt = e / H;
        const int i = (int)(bt % T), b = (int)(bt / T);
        const uint4* po = reinterpret_cast<const uint4*>(o + bt * ldo + h * DK);
        const uint4* pd = reinterpret_cast<const uint4*>(dout + bt * lddo + h * DK);
        float s = 0.f;
#pragma unroll
        for (int u = 0; u < DK / 8; ++u) {
            const uint4 x = po[u], y = pd[u];
            const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&x);
            const __nv_bfloat162* yh = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const float2 xf = __bfloat1622float2(xh[w]), yf = __bfloat1622float2(yh[w]);
                s = fmaf(xf.x, yf.x, s); s = fmaf(xf.y, yf.y, s);
            }
        }
        dsum[((int64_t)b * H + h) * T + i] = s;
    }
}

// dq = bf16(dq_acc + W_full R),  dR += W_full^T q   with W_full[clamp] = -sum_{d<clamp} W[d]; one warp per (b, i, h) row
__global__ void __launch_bounds__(256) attn_bwd_finish_kernel(const float* __restrict__ dq_acc, const float* __restrict__ w,
                                                              const __nv_bfloat16* __restrict__ q, int64_t ldq,
                                                              const __nv_bfloat16* __restrict__ r, int64_t ldr,
                                                              __nv_bfloat16* __restrict__ dq, int64_t lddq,
                                                              float* __restrict__ dr_part,
                                                              int B, int H, int T, int has_rel, int clamp) {
    pdl_entry();
    // Round 1 walked the CTA's 64 rows one per warp iteration, each a dependent chain of loads (accumulator, q, 16 scalar
    // band weights) and finished with shared-memory atomics: 68 us per launch on average for a 12 us data volume
    // (profiles/r02_launches_train_step.md).  Here a warp issues the loads of FOUR rows before it touches any of them (the band
    // weights arrive as ONE 64-byte load, one value per lane, and are broadcast by shuffles), and the per-warp dR partials go
    // through plain stores.
    __shared__ float sdr[8][16][DK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // grid.x = (b, h, row block of 64): all rows of a CTA share the head so that dR reduces in shared memory first
    const int rblocks = (T + 63) / 64;
    const int rb = blockIdx.x % rblocks;
    const int h = (blockIdx.x / rblocks) % H;
    const int b = blockIdx.x / (rblocks * H);
    float rloc[16][2];
    if (has_rel) {
#pragma unroll
        for (int d = 0; d < 16; ++d) {
            if (d <= clamp) {
                const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(r + (int64_t)d * ldr + h * DK + lane * 2);
                const float2 f = __bfloat1622float2(v);
                rloc[d][0] = f.x; rloc[d][1] = f.y;
            } else { rloc[d][0] = 0.f; rloc[d][1] = 0.f; }
        }
    }
    float dracc[16][2];
#pragma unroll
    for (int d = 0; d < 16; ++d) { dracc[d][0] = 0.f; dracc[d][1] = 0.f; }
    constexpr int RB = 4;                                    // rows in flight per warp
#pragma unroll 1
    for (int i0 = warp * 8; i0 < warp * 8 + 8; i0 += RB) {   // warp w owns rows 8w .. 8w+7 of the block
        float2 acc[RB], qf[RB];
        float wl[RB];
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int i = rb * 64 + i0 + u;
            acc[u] = make_float2(0.f, 0.f); qf[u] = make_float2(0.f, 0.f); wl[u] = 0.f;
            if (i < T) {
                const int64_t bt = (int64_t)b * T + i;
                acc[u] = *reinterpret_cast<const float2*>(dq_acc + bt * ((int64_t)H * DK) + h * DK + lane * 2);
                if (has_rel) {
                    qf[u] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(q + bt * ldq + h * DK + lane * 2));
                    if (lane < 16) wl[u] = w[(((int64_t)b * H + h) * T + i) * 16 + lane];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int i = rb * 64 + i0 + u;
            float o0 = acc[u].x, o1 = acc[u].y;
            if (has_rel) {
                float wsum = 0.f;
#pragma unroll
                for (int d = 0; d < 16; ++d) {
                    float wv = __shfl_sync(0xffffffffu, wl[u], d);
                    if (d < clamp) wsum += wv;
                    else if (d == clamp) wv = -wsum;
                    else wv = 0.f;
                    o0 = fmaf(wv, rloc[d][0], o0); o1 = fmaf(wv, rloc[d][1], o1);
                    dracc[d][0] = fmaf(wv, qf[u].x, dracc[d][0]); dracc[d][1] = fmaf(wv, qf[u].y, dracc[d][1]);
                }
            }
            if (i < T)
                *reinterpret_cast<__nv_bfloat162*>(dq + ((int64_t)b * T + i) * lddq + h * DK + lane * 2) = __floats2bfloat162_rn(o0, o1);
        }
    }
    if (has_rel && dr_part) {
#pragma unroll
        for (int d = 0; d < 16; ++d)
            *reinterpret_cast<float2*>(&sdr[warp][d][lane * 2]) = make_float2(dracc[d][0], dracc[d][1]);
        __syncthreads();
        // this CTA's partial [16][64] goes to scratch (plain stores); attn_bwd_dr_reduce_kernel sums the partials of a head:
        // atomics straight into dr would put B*T/64 adds on each of only H*(clamp+1)*64 addresses
        float* dst = dr_part + (int64_t)blockIdx.x * 16 * DK;
        for (int e = threadIdx.x; e < 16 * DK; e += 256) {
            float t = 0.f;
#pragma unroll
            for (int w_ = 0; w_ < 8; ++w_) t += (&sdr[w_][0][0])[e];
            dst[e] = t;
        }
    }
}

// dr[d][h*64 + c] += sum over the (b, row block) partials of head h.  grid = (H * (clamp + 1), slices), 256 threads:
// thread = (column c, one of 4 interleaved partial streams); every address receives `slices` atomic adds.
__global__ void __launch_bounds__(256) attn_bwd_dr_reduce_kernel(const float* __restrict__ dr_part, float* __restrict__ dr,
                                                                 int64_t lddr, int B, int H, int rblocks) {
    pdl_entry();
    __shared__ float part[4][DK];
    const int h = blockIdx.x % H, d = blockIdx.x / H, c = threadIdx.x & 63, stream = threadIdx.x >> 6;
    const int np = B * rblocks;                                   // partials of this head: index p = b * rblocks + rb
    float s = 0.f;
    for (int p = blockIdx.y * 4 + stream; p < np; p += gridDim.y * 4) {
        const int b = p / rblocks, rb = p % rblocks;
        s += dr_part[((((int64_t)b * H + h) * rblocks + rb) * 16 + d) * DK + c];
    }
    part[stream][c] = s;
    __syncthreads();
    if (stream == 0) atomicAdd(dr + (int64_t)d * lddr + h * DK + c, part[0][c] + part[1][c] + part[2][c] + part[3][c]);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool bwd_map3(EncodeTiledFn enc, CUtensorMap* m, const void* base, int64_t ld, int T, int B, int cols, const char* what) {
    if (((uintptr_t)base % 16) != 0 || (ld * 2) % 16 != 0) {
        set_error("attention_bwd_tc: %s must be 16-byte aligned with a 16-byte-multiple pitch", what);
        return false;
    }
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("attention_bwd_tc: tensor map (%s) failed: %d", what, (int)r); return false; }
    return true;
}

}  // namespace

bool get_tma_encode(void** fn);

size_t attention_bwd_tc_workspace_bytes(int B, int H, int T) {
    return sizeof(float) * ((size_t)B * T * H * DK + (size_t)B * H * T * 16 + (size_t)B * H * T +
                            (size_t)B * H * ceil_div(T, 64) * 16 * DK);
}

// Returns NSP_ERR_UNSUPPORTED (nothing launched) when the shape is outside the tensor-core envelope.
nsp_status attention_bwd_tc_dispatch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                     const void* r, int64_t ldr, int rlen, const int32_t* klens, const float* stats,
                                     const void* out, int64_t ldo, const void* dout, int64_t lddo,
                                     void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                                     float* dr, int64_t lddr, int B, int H, int T, int dkdim, int clamp_len, int causal,
                                     int lookahead, int chunk_c, int chunk_l, void* workspace, size_t workspace_bytes,
                                     cudaStream_t st) {
    if (dkdim != DK || !stats) return NSP_ERR_UNSUPPORTED;
    if (r && !(clamp_len >= 1 && clamp_len <= 15)) return NSP_ERR_UNSUPPORTED;
    if (lddk % 8 || lddv % 8 || lddq % 2 || ldq % 8 || ((uintptr_t)dk % 16) || ((uintptr_t)dv % 16) || ((uintptr_t)dq % 4))
        return NSP_ERR_UNSUPPORTED;
    if (ldo % 8 || lddo % 8 || ((uintptr_t)out % 16) || ((uintptr_t)dout % 16)) return NSP_ERR_UNSUPPORTED;
    if (workspace_bytes < attention_bwd_tc_workspace_bytes(B, H, T)) { set_error("attention_bwd_tc: workspace too small"); return NSP_ERR_INVALID; }
    void* fnp = nullptr;
    if (!get_tma_encode(&fnp)) return NSP_ERR_CUDA;
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    CUtensorMap mq, mk, mv, mdo, mr;
    const int cols = H * DK;
    if (!bwd_map3(enc, &mq, q, ldq, T, B, cols, "q")) return NSP_ERR_INVALID;
    if (!bwd_map3(enc, &mk, k, ldk, T, B, cols, "k")) return NSP_ERR_INVALID;
    if (!bwd_map3(enc, &mv, v, ldv, T, B, cols, "v")) return NSP_ERR_INVALID;
    if (!bwd_map3(enc, &mdo, dout, lddo, T, B, cols, "dout")) return NSP_ERR_INVALID;
    mr = mq;
    if (r) {
        if (((uintptr_t)r % 16) != 0 || (ldr * 2) % 16 != 0) { set_error("attention_bwd_tc: r alignment"); return NSP_ERR_INVALID; }
        cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rlen};
        cuuint64_t gstr[1] = {(cuuint64_t)ldr * 2};
        cuuint32_t box[2] = {64, 16};
        cuuint32_t es[2] = {1, 1};
        CUresult rc = enc(&mr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(r), gdim, gstr, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) { set_error("attention_bwd_tc: tensor map (r) failed: %d", (int)rc); return NSP_ERR_CUDA; }
    }
    float* dq_acc = (float*)workspace;
    float* w = dq_acc + (size_t)B * T * H * DK;
    float* dsum = w + (size_t)B * H * T * 16;
    NSP_CUDA_OK(cudaMemsetAsync(workspace, 0, sizeof(float) * ((size_t)B * T * H * DK + (size_t)B * H * T * 16), st));
    {
        const int64_t n = (int64_t)B * T * H;
        int grid = (int)((n + 255) / 256);
        launch_k(attn_bwd_prep_kernel, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)out, ldo, (const __nv_bfloat16*)dout, lddo, dsum, B, H, T);
        NSP_LAUNCH_OK();
    }
    BwdTcArgs a;
    a.klens = klens; a.stats = stats; a.dsum = dsum; a.dq_acc = dq_acc; a.w = w;
    a.dk = (__nv_bfloat16*)dk; a.lddk = lddk; a.dv = (__nv_bfloat16*)dv; a.lddv = lddv;
    a.B = B; a.H = H; a.T = T; a.has_rel = r ? 1 : 0;
    a.clamp = r ? (clamp_len < rlen - 1 ? clamp_len : rlen - 1) : 0;
    a.causal = causal; a.lookahead = lookahead; a.chunk_c = chunk_c; a.chunk_l = chunk_l;
    a.inv_scale = 1.0f / sqrtf((float)DK);
    a.scale_log2 = 1.4426950408889634f * a.inv_scale;
    const size_t smem = 1024 + 10 * TILE + 2048 + (128 * 17 + 3 * 128) * sizeof(float) + 128;
    static bool attr = false;
    if (!attr) { NSP_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    const int ktiles = ceil_div(T, KT);
    launch_k(attn_bwd_tc_kernel, dim3((unsigned)(B * H * ktiles)), dim3(NTHREADS), smem, st, mq, mk, mv, mdo, mr, a);
    NSP_LAUNCH_OK();
    float* dr_part = dsum + (size_t)B * H * T;
    const int rblocks = ceil_div(T, 64);
    launch_k(attn_bwd_finish_kernel, dim3((unsigned)(B * H * rblocks)), dim3(256), 0, st, dq_acc, w, (const __nv_bfloat16*)q, ldq,
                                                                      (const __nv_bfloat16*)r, ldr,
                                                                      (__nv_bfloat16*)dq, lddq, (a.has_rel && dr) ? dr_part : nullptr,
                                                                      B, H, T, a.has_rel, a.clamp);
    NSP_LAUNCH_OK();
    if (a.has_rel && dr) {
        const int np = B * rblocks;
        dim3 rgrid((unsigned)(H * (a.clamp + 1)), (unsigned)(np >= 64 ? 8 : 1));
        launch_k(attn_bwd_dr_reduce_kernel, dim3(rgrid), dim3(256), 0, st, dr_part, dr, lddr, B, H, rblocks);
        NSP_LAUNCH_OK();
    }
    return NSP_OK;
}

}  // namespace nsp
