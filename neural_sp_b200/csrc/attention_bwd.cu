// Backward of the (relative-position) multi-head self-attention core, flash-style: the T x T score / weight matrices
// are recomputed tile by tile and never reach HBM.  fp32 math on CUDA cores (bf16 or fp32 I/O).
//
// The reference obtains this from autograd over  RelativeMultiheadAttentionMechanism.forward
//   modules/relative_multihead_attention.py:146-220  ((q+u) k^T + rel_shift((q+v) R^T), /sqrt(d_k), masked_fill, softmax, aw v)
// and MultiheadAttentionMechanism.forward modules/multihead_attention.py:93-157 (r == NULL).
//
//   S_ij = ((q_i+u).k_j + (q_i+v).r[dist(i,j)]) / sqrt(dk),  P = softmax_j(S) (masked keys: finfo.min, no gradient),
//   O = P V;   D_i = dO_i . O_i;   dP = dO V^T;   dS = P * (dP - D) / sqrt(dk)  (0 where masked)
//   dq_i = sum_j dS_ij k_j + sum_d W_id r_d        W_id = sum_{j: dist(i,j)=d} dS_ij
//   dk_j = sum_i dS_ij (q_i+u);   dv_j = sum_i P_ij dO_i;   dr_d = sum_i W_id (q_i+v);   du = sum_i dq^AC_i;  dv_bias = sum_i dq^BD_i
//
// Kernel A (one CTA per 32-query tile): pass 1 recomputes the softmax statistics (m, 1/l) and the table
//   QvR[i][d] = (q_i+v).r_d, pass 2 forms dS and accumulates dq, W; writes dq, (m, 1/l, D) and QvR for kernel B,
//   and adds dr / du / dv_bias with atomics.
// Kernel B (one CTA per 64-key tile): loops over the query tiles, rebuilds P and dS from the saved statistics and
//   accumulates dk, dv.
#include <float.h>
#include "common.cuh"

namespace nsp {
namespace {

constexpr int QT = 32;
constexpr int KT = 64;
constexpr int QP = QT + 4;
constexpr int KP = KT + 4;

struct AttnBwdParams {
    const void* q; const void* k; const void* v; int64_t ldq, ldk, ldv;
    const void* r; int64_t ldr; int rlen;
    const float* u_bias; const float* v_bias;
    const int32_t* klens;
    const void* o; int64_t ldo;
    const void* dout; int64_t lddo;
    void* dqp; void* dkp; void* dvp; int64_t lddq, lddk, lddv;
    float* dr; int64_t lddr;
    float* du; float* dvb;
    float* stats;          // [B*H*Tq][3]  (m, 1/l, D)
    float* qvr;            // [B*H*Tq][ndp]
    int ndp;
    int B, H, Tq, Tk, dk;
    int ceff;              // effective clamp of the distance: min(clamp_len (if > 0), rlen - 1)
    int causal, lookahead, chunk_c, chunk_l;
    float inv_scale;
};

template <typename T> __device__ __forceinline__ float ab_ld(const T* p);
template <> __device__ __forceinline__ float ab_ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ab_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void ab_st(T* p, float v);
template <> __device__ __forceinline__ void ab_st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void ab_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ bool key_visible(const AttnBwdParams& p, int i, int j, int klen, int mlen) {
    bool vis = j < klen;
    if (p.causal) vis = vis && (j <= mlen + i + p.lookahead);
    if (p.chunk_c > 0) {
        const int cs = ((mlen + i) / p.chunk_c) * p.chunk_c;
        vis = vis && (j >= cs - p.chunk_l) && (j < cs + p.chunk_c);
    }
    return vis;
}
__device__ __forceinline__ int rel_dist(int x, int ceff) { x = x < 0 ? -x : x; return x < ceff ? x : ceff; }

// C[4][4] = sum_c AT[c][ty*4+e] * BT[c][tx*4+f]   (both operands transposed in smem, row pitches QP / KP)
template <int DKP>
__device__ __forceinline__ void tile_qk(const float* __restrict__ AT, const float* __restrict__ BT, float (&C)[4][4], int ty, int tx) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int f = 0; f < 4; ++f) C[e][f] = 0.f;
#pragma unroll 8
    for (int c = 0; c < DKP; ++c) {
        const float4 a = *reinterpret_cast<const float4*>(AT + c * QP + ty * 4);
        const float4 b = *reinterpret_cast<const float4*>(BT + c * KP + tx * 4);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int f = 0; f < 4; ++f) C[e][f] = fmaf(av[e], bv[f], C[e][f]);
    }
}

// scores of one tile: S = (AC + BD) * inv_scale with the reference's masking (-FLT_MAX masked, -INF beyond the tensor)
template <int DKP>
__device__ __forceinline__ void tile_scores(const AttnBwdParams& p, const float* QuT, const float* KsT, const float* qvr_rows,
                                            float (&S)[4][4], bool (&vis)[4][4], int ty, int tx, int i0, int j0, int klen, int mlen) {
    tile_qk<DKP>(QuT, KsT, S, ty, tx);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int i = i0 + ty * 4 + e;
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            const int j = j0 + tx * 4 + f;
            float s = S[e][f];
            if (qvr_rows && i < p.Tq) s += qvr_rows[(int64_t)(ty * 4 + e) * p.ndp + rel_dist(mlen + i - j, p.ceff)];
            s *= p.inv_scale;
            const bool vv = key_visible(p, i, j, klen, mlen);
            vis[e][f] = vv && j < p.Tk;
            if (!vv) s = -FLT_MAX;
            if (j >= p.Tk) s = -INFINITY;
            S[e][f] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel A
// ------------------------------------------------------------------------------------------------
template <typename T, int DKP>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(AttnBwdParams p) {
    pdl_entry();
    constexpr int CPT = DKP / 16;
    extern __shared__ float sm[];
    float* QuT = sm;                              // [DKP][QP]
    float* QvT = QuT + DKP * QP;                  // [DKP][QP]
    float* dOT = QvT + DKP * QP;                  // [DKP][QP]
    float* KsT = dOT + DKP * QP;                  // [DKP][KP]   (prologue: O^T)
    float* VsT = KsT + DKP * KP;                  // [DKP][KP]
    float* Ks = VsT + DKP * KP;                   // [KT][DKP+4]
    float* dSsT = Ks + KT * (DKP + 4);            // [KT][QP]
    float* Ds = dSsT + KT * QP;                   // [QT]
    float* Wt = Ds + QT;                          // [QT][ndp]

    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int qtiles = (p.Tq + QT - 1) / QT;
    const int qt = blockIdx.x % qtiles;
    const int h = (blockIdx.x / qtiles) % p.H;
    const int b = blockIdx.x / (qtiles * p.H);
    const int i0 = qt * QT;
    const int dk = p.dk;
    const int mlen = p.Tk - p.Tq;
    const int klen = min(max(p.klens[b], 0), p.Tk);
    const T* qg = reinterpret_cast<const T*>(p.q) + (int64_t)b * p.Tq * p.ldq + (int64_t)h * dk;
    const T* kg = reinterpret_cast<const T*>(p.k) + (int64_t)b * p.Tk * p.ldk + (int64_t)h * dk;
    const T* vg = reinterpret_cast<const T*>(p.v) + (int64_t)b * p.Tk * p.ldv + (int64_t)h * dk;
    const T* og = reinterpret_cast<const T*>(p.o) + (int64_t)b * p.Tq * p.ldo + (int64_t)h * dk;
    const T* dog = reinterpret_cast<const T*>(p.dout) + (int64_t)b * p.Tq * p.lddo + (int64_t)h * dk;
    const T* rg = p.r ? reinterpret_cast<const T*>(p.r) + (int64_t)h * dk : nullptr;
    const int64_t row0 = ((int64_t)b * p.H + h) * p.Tq + i0;        // first row of this tile in stats / qvr
    float* qvr_rows = rg ? p.qvr + row0 * p.ndp : nullptr;

    // ---- stage Q(+u,+v)^T, dO^T, O^T ----
    for (int e = tid; e < QT * DKP; e += 128) {
        const int qi = e % QT, c = e / QT;
        float val = 0.f, dov = 0.f, ov = 0.f;
        if (c < dk && i0 + qi < p.Tq) {
            val = ab_ld<T>(qg + (int64_t)(i0 + qi) * p.ldq + c);
            dov = ab_ld<T>(dog + (int64_t)(i0 + qi) * p.lddo + c);
            ov = ab_ld<T>(og + (int64_t)(i0 + qi) * p.ldo + c);
        }
        const float ub = (p.u_bias && c < dk) ? p.u_bias[h * dk + c] : 0.f;
        const float vb = (p.v_bias && c < dk) ? p.v_bias[h * dk + c] : 0.f;
        QuT[c * QP + qi] = val + ub;
        QvT[c * QP + qi] = val + vb;
        dOT[c * QP + qi] = dov;
        KsT[c * QP + qi] = ov;            // scratch: O^T with pitch QP
    }
    for (int e = tid; e < QT * p.ndp; e += 128) Wt[e] = 0.f;
    __syncthreads();
    if (tid < QT) {
        float s = 0.f;
        for (int c = 0; c < DKP; ++c) s = fmaf(dOT[c * QP + tid], KsT[c * QP + tid], s);
        Ds[tid] = s;
    }
    if (rg) {
        const int nd = p.ceff + 1;
        for (int e = tid; e < QT * nd; e += 128) {
            const int qi = e % QT, d = e / QT;
            float s = 0.f;
            for (int c = 0; c < dk; ++c) s = fmaf(QvT[c * QP + qi], ab_ld<T>(rg + (int64_t)d * p.ldr + c), s);
            if (i0 + qi < p.Tq) qvr_rows[(int64_t)qi * p.ndp + d] = s;
        }
    }
    __syncthreads();

    // ---- pass 1: softmax statistics ----
    float m_run[4], l_run[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { m_run[e] = -FLT_MAX; l_run[e] = 0.f; }
    for (int j0 = 0; j0 < p.Tk; j0 += KT) {
        __syncthreads();
        for (int e = tid; e < KT * DKP; e += 128) {
            const int kj = e % KT, c = e / KT;
            float val = 0.f;
            if (c < dk && j0 + kj < p.Tk) val = ab_ld<T>(kg + (int64_t)(j0 + kj) * p.ldk + c);
            KsT[c * KP + kj] = val;
        }
        __syncthreads();
        float S[4][4]; bool vis[4][4];
        tile_scores<DKP>(p, QuT, KsT, qvr_rows, S, vis, ty, tx, i0, j0, klen, mlen);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float rmax = fmaxf(fmaxf(S[e][0], S[e][1]), fmaxf(S[e][2], S[e][3]));
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
            const float m_new = fmaxf(m_run[e], rmax);
            float rsum = 0.f;
#pragma unroll
            for (int f = 0; f < 4; ++f) rsum += (S[e][f] == -INFINITY) ? 0.f : __expf(S[e][f] - m_new);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
            l_run[e] = l_run[e] * __expf(m_run[e] - m_new) + rsum;
            m_run[e] = m_new;
        }
    }
    float inv_l[4], Dq[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        inv_l[e] = 1.f / l_run[e];
        Dq[e] = Ds[ty * 4 + e];
        const int i = i0 + ty * 4 + e;
        if (tx == 0 && i < p.Tq) {
            float* st = p.stats + (row0 + ty * 4 + e) * 3;
            st[0] = m_run[e]; st[1] = inv_l[e]; st[2] = Dq[e];
        }
    }

    // ---- pass 2: dS, dq, W ----
    float dQ[4][CPT];
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int f = 0; f < CPT; ++f) dQ[e][f] = 0.f;
    for (int j0 = 0; j0 < p.Tk; j0 += KT) {
        __syncthreads();
        for (int e = tid; e < KT * DKP; e += 128) {
            const int kj = e % KT, c = e / KT;
            float kv = 0.f, vv = 0.f;
            if (c < dk && j0 + kj < p.Tk) {
                kv = ab_ld<T>(kg + (int64_t)(j0 + kj) * p.ldk + c);
                vv = ab_ld<T>(vg + (int64_t)(j0 + kj) * p.ldv + c);
            }
            KsT[c * KP + kj] = kv;
            VsT[c * KP + kj] = vv;
            Ks[kj * (DKP + 4) + c] = kv;
        }
        __syncthreads();
        float S[4][4], dP[4][4]; bool vis[4][4];
        tile_scores<DKP>(p, QuT, KsT, qvr_rows, S, vis, ty, tx, i0, j0, klen, mlen);
        tile_qk<DKP>(dOT, VsT, dP, ty, tx);
        // uniform saturation test of the relative distance over the whole tile
        const int lo = mlen + i0 - j0 - (KT - 1), hi = mlen + i0 + (QT - 1) - j0;
        const bool saturated = rg && (lo >= p.ceff || hi <= -p.ceff);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = i0 + ty * 4 + e;
            float rowsum = 0.f;
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const float pv = (S[e][f] == -INFINITY) ? 0.f : __expf(S[e][f] - m_run[e]) * inv_l[e];
                const float ds = vis[e][f] ? pv * (dP[e][f] - Dq[e]) * p.inv_scale : 0.f;
                S[e][f] = ds;
                rowsum += ds;
                if (rg && !saturated && ds != 0.f)
                    atomicAdd(Wt + (ty * 4 + e) * p.ndp + rel_dist(mlen + i - (j0 + tx * 4 + f), p.ceff), ds);
            }
            if (saturated) {
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) rowsum += __shfl_xor_sync(0xffffffffu, rowsum, o);
                if (tx == 0) Wt[(ty * 4 + e) * p.ndp + p.ceff] += rowsum;     // this (ty, e) row is owned by one thread here
            }
        }
#pragma unroll
        for (int f = 0; f < 4; ++f)
            *reinterpret_cast<float4*>(dSsT + (tx * 4 + f) * QP + ty * 4) = make_float4(S[0][f], S[1][f], S[2][f], S[3][f]);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < KT; ++j) {
            const float4 a = *reinterpret_cast<const float4*>(dSsT + j * QP + ty * 4);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float* kr = Ks + j * (DKP + 4) + tx * CPT;
#pragma unroll
            for (int f = 0; f < CPT; ++f) {
                const float kv = kr[f];
#pragma unroll
                for (int e = 0; e < 4; ++e) dQ[e][f] = fmaf(av[e], kv, dQ[e][f]);
            }
        }
    }
    __syncthreads();

    // ---- epilogue: dq = dq^AC + W r ; dr, du, dv_bias ----
    float dQb[4][CPT];
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int f = 0; f < CPT; ++f) dQb[e][f] = 0.f;
    if (rg) {
        const int nd = p.ceff + 1;
        for (int d = 0; d < nd; ++d) {
            float rv[CPT];
#pragma unroll
            for (int f = 0; f < CPT; ++f) { const int c = tx * CPT + f; rv[f] = c < dk ? ab_ld<T>(rg + (int64_t)d * p.ldr + c) : 0.f; }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float w = Wt[(ty * 4 + e) * p.ndp + d];
#pragma unroll
                for (int f = 0; f < CPT; ++f) dQb[e][f] = fmaf(w, rv[f], dQb[e][f]);
            }
        }
        if (p.dr) {
            for (int e = tid; e < nd * dk; e += 128) {
                const int c = e % dk, d = e / dk;
                float s = 0.f;
#pragma unroll 8
                for (int qi = 0; qi < QT; ++qi) s = fmaf(Wt[qi * p.ndp + d], QvT[c * QP + qi], s);
                atomicAdd(p.dr + (int64_t)d * p.lddr + h * dk + c, s);
            }
        }
    }
    T* dqg = reinterpret_cast<T*>(p.dqp) + (int64_t)b * p.Tq * p.lddq + (int64_t)h * dk;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int i = i0 + ty * 4 + e;
        if (i < p.Tq) {
#pragma unroll
            for (int f = 0; f < CPT; ++f) {
                const int c = tx * CPT + f;
                if (c < dk) ab_st<T>(dqg + (int64_t)i * p.lddq + c, dQ[e][f] + dQb[e][f]);
            }
        }
    }
    if (p.du || p.dvb) {
        float* red = KsT;          // [2][DKP] scratch (all tile reads are done: barrier above)
        for (int e = tid; e < 2 * DKP; e += 128) red[e] = 0.f;
        __syncthreads();
#pragma unroll
        for (int f = 0; f < CPT; ++f) {
            float sa = 0.f, sb = 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) { sa += dQ[e][f]; sb += dQb[e][f]; }     // rows beyond Tq are exactly zero
            atomicAdd(red + tx * CPT + f, sa);
            atomicAdd(red + DKP + tx * CPT + f, sb);
        }
        __syncthreads();
        for (int c = tid; c < dk; c += 128) {
            if (p.du) atomicAdd(p.du + h * dk + c, red[c]);
            if (p.dvb) atomicAdd(p.dvb + h * dk + c, red[DKP + c]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel B
// ------------------------------------------------------------------------------------------------
template <typename T, int DKP>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(AttnBwdParams p) {
    pdl_entry();
    constexpr int CPB = DKP / 8;                  // output columns per thread
    extern __shared__ float sm[];
    float* KsT = sm;                              // [DKP][KP]
    float* VsT = KsT + DKP * KP;                  // [DKP][KP]
    float* QuT = VsT + DKP * KP;                  // [DKP][QP]
    float* dOT = QuT + DKP * QP;                  // [DKP][QP]
    float* Qu = dOT + DKP * QP;                   // [QT][DKP+4]
    float* dOr = Qu + QT * (DKP + 4);             // [QT][DKP+4]
    float* Ps = dOr + QT * (DKP + 4);             // [QT][KP]
    float* dSs = Ps + QT * KP;                    // [QT][KP]
    float* St = dSs + QT * KP;                    // [3][QT]

    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int ky = tid >> 3, cx = tid & 7;
    const int ktiles = (p.Tk + KT - 1) / KT;
    const int kt = blockIdx.x % ktiles;
    const int h = (blockIdx.x / ktiles) % p.H;
    const int b = blockIdx.x / (ktiles * p.H);
    const int j0 = kt * KT;
    const int dk = p.dk;
    const int mlen = p.Tk - p.Tq;
    const int klen = min(max(p.klens[b], 0), p.Tk);
    const T* qg = reinterpret_cast<const T*>(p.q) + (int64_t)b * p.Tq * p.ldq + (int64_t)h * dk;
    const T* kg = reinterpret_cast<const T*>(p.k) + (int64_t)b * p.Tk * p.ldk + (int64_t)h * dk;
    const T* vg = reinterpret_cast<const T*>(p.v) + (int64_t)b * p.Tk * p.ldv + (int64_t)h * dk;
    const T* dog = reinterpret_cast<const T*>(p.dout) + (int64_t)b * p.Tq * p.lddo + (int64_t)h * dk;
    const bool has_r = p.r != nullptr;

    for (int e = tid; e < KT * DKP; e += 128) {
        const int kj = e % KT, c = e / KT;
        float kv = 0.f, vv = 0.f;
        if (c < dk && j0 + kj < p.Tk) {
            kv = ab_ld<T>(kg + (int64_t)(j0 + kj) * p.ldk + c);
            vv = ab_ld<T>(vg + (int64_t)(j0 + kj) * p.ldv + c);
        }
        KsT[c * KP + kj] = kv;
        VsT[c * KP + kj] = vv;
    }
    float dK[4][CPB], dV[4][CPB];
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int f = 0; f < CPB; ++f) { dK[e][f] = 0.f; dV[e][f] = 0.f; }

    for (int i0 = 0; i0 < p.Tq; i0 += QT) {
        const int64_t row0 = ((int64_t)b * p.H + h) * p.Tq + i0;
        __syncthreads();
        for (int e = tid; e < QT * DKP; e += 128) {
            const int c = e % DKP, qi = e / DKP;
            float val = 0.f, dov = 0.f;
            if (c < dk && i0 + qi < p.Tq) {
                val = ab_ld<T>(qg + (int64_t)(i0 + qi) * p.ldq + c) + (p.u_bias ? p.u_bias[h * dk + c] : 0.f);
                dov = ab_ld<T>(dog + (int64_t)(i0 + qi) * p.lddo + c);
            }
            QuT[c * QP + qi] = val;
            dOT[c * QP + qi] = dov;
            Qu[qi * (DKP + 4) + c] = val;
            dOr[qi * (DKP + 4) + c] = dov;
        }
        if (tid < QT) {
            const bool ok = i0 + tid < p.Tq;
            const float* st = p.stats + (row0 + tid) * 3;
            St[tid] = ok ? st[0] : 0.f;
            St[QT + tid] = ok ? st[1] : 0.f;
            St[2 * QT + tid] = ok ? st[2] : 0.f;
        }
        __syncthreads();
        float S[4][4], dP[4][4]; bool vis[4][4];
        tile_qk<DKP>(QuT, KsT, S, ty, tx);
        tile_qk<DKP>(dOT, VsT, dP, ty, tx);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int qi = ty * 4 + e, i = i0 + qi;
            const bool row_ok = i < p.Tq;
            const float m = St[qi], il = St[QT + qi], Dq = St[2 * QT + qi];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const int j = j0 + tx * 4 + f;
                float s = S[e][f];
                if (has_r && row_ok) s += p.qvr[(row0 + qi) * p.ndp + rel_dist(mlen + i - j, p.ceff)];
                s *= p.inv_scale;
                const bool vv = key_visible(p, i, j, klen, mlen);
                if (!vv) s = -FLT_MAX;
                const float pv = (j >= p.Tk || !row_ok) ? 0.f : __expf(s - m) * il;
                const float ds = (vv && j < p.Tk) ? pv * (dP[e][f] - Dq) * p.inv_scale : 0.f;
                Ps[qi * KP + tx * 4 + f] = pv;
                dSs[qi * KP + tx * 4 + f] = ds;
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int qi = 0; qi < QT; ++qi) {
            const float4 pa = *reinterpret_cast<const float4*>(Ps + qi * KP + ky * 4);
            const float4 da = *reinterpret_cast<const float4*>(dSs + qi * KP + ky * 4);
            const float pv[4] = {pa.x, pa.y, pa.z, pa.w}, dv4[4] = {da.x, da.y, da.z, da.w};
            const float* dor = dOr + qi * (DKP + 4) + cx * CPB;
            const float* qur = Qu + qi * (DKP + 4) + cx * CPB;
#pragma unroll
            for (int f = 0; f < CPB; ++f) {
                const float dov = dor[f], quv = qur[f];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    dV[e][f] = fmaf(pv[e], dov, dV[e][f]);
                    dK[e][f] = fmaf(dv4[e], quv, dK[e][f]);
                }
            }
        }
    }
    T* dkg = reinterpret_cast<T*>(p.dkp) + (int64_t)b * p.Tk * p.lddk + (int64_t)h * dk;
    T* dvg = reinterpret_cast<T*>(p.dvp) + (int64_t)b * p.Tk * p.lddv + (int64_t)h * dk;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int j = j0 + ky * 4 + e;
        if (j < p.Tk) {
#pragma unroll
            for (int f = 0; f < CPB; ++f) {
                const int c = cx * CPB + f;
                if (c < dk) {
                    ab_st<T>(dkg + (int64_t)j * p.lddk + c, dK[e][f]);
                    ab_st<T>(dvg + (int64_t)j * p.lddv + c, dV[e][f]);
                }
            }
        }
    }
}

template <typename T, int DKP>
nsp_status launch_bwd(const AttnBwdParams& p, cudaStream_t st) {
    const size_t smem_a = sizeof(float) * ((size_t)3 * DKP * QP + (size_t)2 * DKP * KP + (size_t)KT * (DKP + 4) +
                                           (size_t)KT * QP + QT + (size_t)QT * p.ndp);
    const size_t smem_b = sizeof(float) * ((size_t)2 * DKP * KP + (size_t)2 * DKP * QP + (size_t)2 * QT * (DKP + 4) +
                                           (size_t)2 * QT * KP + 3 * QT);
    if (smem_a > 225 * 1024) {
        set_error("attention_bwd: position table with %d rows at d_k=%d needs %zu B of shared memory", p.ceff + 1, p.dk, smem_a);
        return NSP_ERR_UNSUPPORTED;
    }
    auto ka = attn_bwd_dq_kernel<T, DKP>;
    auto kb = attn_bwd_dkv_kernel<T, DKP>;
    static size_t attr_a = 0, attr_b = 0;
    if (smem_a > attr_a) { NSP_CUDA_OK(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a)); attr_a = smem_a; }
    if (smem_b > attr_b) { NSP_CUDA_OK(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b)); attr_b = smem_b; }
    launch_k(ka, dim3((unsigned)(p.B * p.H * ceil_div(p.Tq, QT))), dim3(128), smem_a, st, p);
    NSP_LAUNCH_OK();
    launch_k(kb, dim3((unsigned)(p.B * p.H * ceil_div(p.Tk, KT))), dim3(128), smem_b, st, p);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

}  // namespace
}  // namespace nsp

namespace nsp {
size_t attention_bwd_tc_workspace_bytes(int B, int H, int T);
nsp_status attention_bwd_tc_dispatch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                     const void* r, int64_t ldr, int rlen, const int32_t* klens, const float* stats,
                                     const void* out, int64_t ldo, const void* dout, int64_t lddo,
                                     void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                                     float* dr, int64_t lddr, int B, int H, int T, int dkdim, int clamp_len, int causal,
                                     int lookahead, int chunk_c, int chunk_l, void* workspace, size_t workspace_bytes,
                                     cudaStream_t st);
}

using namespace nsp;

static int attn_bwd_ndp(int rlen, int clamp_len, int has_r) {
    if (!has_r) return 4;
    int ceff = rlen - 1;
    if (clamp_len > 0 && clamp_len < ceff) ceff = clamp_len;
    return (ceff + 1 + 3) / 4 * 4;
}

extern "C" size_t nsp_relpos_attention_bwd_workspace_bytes(int B, int H, int Tq, int rlen, int clamp_len, int has_r) {
    const size_t rows = (size_t)B * H * Tq;
    const size_t simt = sizeof(float) * rows * (3 + (size_t)attn_bwd_ndp(rlen, clamp_len, has_r));
    const size_t tcb = attention_bwd_tc_workspace_bytes(B, H, Tq);
    return simt > tcb ? simt : tcb;
}

extern "C" nsp_status nsp_relpos_attention_bwd(int is_bf16, const void* q, int64_t ldq, const void* k, int64_t ldk,
                                               const void* v, int64_t ldv, const void* r, int64_t ldr, int rlen,
                                               const float* u_bias, const float* v_bias, const int32_t* klens,
                                               const void* out, int64_t ldo, const void* dout, int64_t lddo,
                                               void* dq, int64_t lddq, void* dk_, int64_t lddk, void* dv, int64_t lddv,
                                               float* dr, int64_t lddr, float* du, float* dvb, const float* stats,
                                               int B, int H, int Tq, int Tk, int dk, int clamp_len, int causal, int lookahead,
                                               int chunk_c, int chunk_l, void* workspace, size_t workspace_bytes, void* stream) {
    NSP_CHECK_ARG(q && k && v && klens && out && dout && dq && dk_ && dv && workspace, "attention_bwd: null pointer");
    NSP_CHECK_ARG(B > 0 && H > 0 && Tq > 0 && Tk >= Tq && dk > 0, "attention_bwd: bad shape B=%d H=%d Tq=%d Tk=%d dk=%d", B, H, Tq, Tk, dk);
    NSP_CHECK_ARG(!r || rlen > 0, "attention_bwd: rlen must be positive when r is given");
    if (dk > 128) { set_error("attention_bwd: d_k=%d unsupported (max 128)", dk); return NSP_ERR_UNSUPPORTED; }
    NSP_CHECK_ARG(workspace_bytes >= nsp_relpos_attention_bwd_workspace_bytes(B, H, Tq, rlen, clamp_len, r != nullptr),
                  "attention_bwd: workspace too small");
    if (is_bf16 && stats && !u_bias && !v_bias && Tq == Tk) {
        // tensor-core path (needs the forward kernel's softmax statistics); falls through when outside its envelope
        nsp_status s = attention_bwd_tc_dispatch(q, ldq, k, ldk, v, ldv, r, ldr, rlen, klens, stats, out, ldo, dout, lddo,
                                                 dq, lddq, dk_, lddk, dv, lddv, dr, lddr, B, H, Tq, dk, clamp_len, causal,
                                                 lookahead, chunk_c, chunk_l, workspace, workspace_bytes, (cudaStream_t)stream);
        if (s != NSP_ERR_UNSUPPORTED) return s;
    }
    AttnBwdParams p;
    p.q = q; p.k = k; p.v = v; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.r = r; p.ldr = ldr; p.rlen = rlen;
    p.u_bias = u_bias; p.v_bias = v_bias; p.klens = klens; p.o = out; p.ldo = ldo; p.dout = dout; p.lddo = lddo;
    p.dqp = dq; p.dkp = dk_; p.dvp = dv; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv; p.dr = dr; p.lddr = lddr;
    p.du = du; p.dvb = dvb;
    p.B = B; p.H = H; p.Tq = Tq; p.Tk = Tk; p.dk = dk;
    p.ceff = 0;
    if (r) { p.ceff = rlen - 1; if (clamp_len > 0 && clamp_len < p.ceff) p.ceff = clamp_len; }
    p.ndp = attn_bwd_ndp(rlen, clamp_len, r != nullptr);
    p.causal = causal; p.lookahead = lookahead; p.chunk_c = chunk_c; p.chunk_l = chunk_l;
    p.inv_scale = 1.0f / sqrtf((float)dk);
    p.stats = (float*)workspace;
    p.qvr = p.stats + (size_t)B * H * Tq * 3;
    cudaStream_t st = (cudaStream_t)stream;
    if (is_bf16) {
        if (dk <= 16) return launch_bwd<__nv_bfloat16, 16>(p, st);
        if (dk <= 64) return launch_bwd<__nv_bfloat16, 64>(p, st);
        return launch_bwd<__nv_bfloat16, 128>(p, st);
    }
    if (dk <= 16) return launch_bwd<float, 16>(p, st);
    if (dk <= 64) return launch_bwd<float, 64>(p, st);
    return launch_bwd<float, 128>(p, st);
}
