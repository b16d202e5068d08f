// Relative-position multi-head self-attention, flash-style (scores never touch HBM), exact fp32 math
// on CUDA cores.  This is the parity-mode kernel (and the fallback shape-generic one); the bf16
// tensor-core kernel lives in attention_tc.cu.
//
// Replaces  RelativeMultiheadAttentionMechanism.forward  modules/relative_multihead_attention.py:146-220
//           (AC = (q+u) k^T, BD = rel_shift((q+v) R^T), e = (AC+BD)/sqrt(d_k), masked_fill(finfo.min),
//            softmax over keys, cv = aw v)  and  _rel_shift :112-144  (dist = |mlen + i - j| clamped),
//           MultiheadAttentionMechanism.forward  modules/multihead_attention.py:93-157 (R == NULL),
//           make_san_mask / causal / make_chunkwise_san_mask  encoders/transformer.py:633-686.
//
// One CTA (128 threads) owns a 32-query tile of one (batch, head) and walks the keys in tiles of 64:
// S = Qu K^T and BD = Qv R_band^T are 4x4 / 4x6 register-tiled fp32 GEMMs out of transposed smem,
// the band of BD is gathered by (i - j), online softmax keeps (m, l) per query row, O accumulates P V.
#include <float.h>
#include "common.cuh"

namespace nsp {
namespace {

constexpr int QT = 32;     // queries per CTA
constexpr int KT = 64;     // keys per tile
constexpr int OB = 96;     // distinct (i-j) offsets in a tile: QT + KT - 1 = 95, padded
constexpr int QP = QT + 4; // padded row of transposed Q / P
constexpr int KP = KT + 4;
constexpr int OP = OB + 4;
constexpr int BP = OB + 1;

struct AttnParams {
    const void* q; const void* k; const void* v;   // element [b*T + t][h*dk + c] with row pitches ldq/ldk/ldv
    int64_t ldq, ldk, ldv;
    const void* r; int64_t ldr; int rlen;          // projected positions [rlen][h*dk + c] or null
    const float* u_bias; const float* v_bias;      // [H, dk] or null
    const int32_t* klens;                          // [B] valid keys per utterance (already including cache)
    void* out; int64_t ldo;
    int B, H, Tq, Tk, dk;
    int clamp_len;                                 // <= 0: no clamp
    int causal, lookahead;                         // causal: key j visible iff j <= mlen + i + lookahead
    int chunk_c, chunk_l;                          // chunk-wise mask (0: off): keys in [chunk_start - chunk_l, chunk_end)
    float inv_scale;
};

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T, int DKP>
__global__ void __launch_bounds__(128) attn_simt_kernel(AttnParams p) {
    pdl_entry();
    constexpr int CPT = DKP / 16;   // output columns per thread
    extern __shared__ float sm[];
    float* QuT = sm;                          // [DKP][QP]
    float* QvT = QuT + DKP * QP;              // [DKP][QP]
    float* KsT = QvT + DKP * QP;              // [DKP][KP]
    float* Vs  = KsT + DKP * KP;              // [KT][DKP+4]
    float* RsT = Vs + KT * (DKP + 4);         // [DKP][OP]
    float* BDs = RsT + DKP * OP;              // [QT][BP]
    float* PsT = BDs + QT * BP;               // [KT][QP]

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
    const T* rg = p.r ? reinterpret_cast<const T*>(p.r) + (int64_t)h * dk : nullptr;

    // ---- stage Q (+u, +v), transposed ----
    for (int e = tid; e < QT * DKP; e += 128) {
        int qi = e % QT, c = e / QT;
        float val = 0.f;
        if (c < dk && i0 + qi < p.Tq) val = ldf<T>(qg + (int64_t)(i0 + qi) * p.ldq + c);
        float ub = (p.u_bias && c < dk) ? p.u_bias[h * dk + c] : 0.f;
        float vb = (p.v_bias && c < dk) ? p.v_bias[h * dk + c] : 0.f;
        QuT[c * QP + qi] = val + ub;
        QvT[c * QP + qi] = val + vb;
    }

    float m_run[4], l_run[4], O[4][CPT];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        m_run[e] = -FLT_MAX; l_run[e] = 0.f;
#pragma unroll
        for (int f = 0; f < CPT; ++f) O[e][f] = 0.f;
    }

    for (int j0 = 0; j0 < p.Tk; j0 += KT) {
        __syncthreads();   // previous tile's smem fully consumed (and Q staged, first time)
        // ---- stage K^T, V, R band^T ----
        for (int e = tid; e < KT * DKP; e += 128) {
            int kj = e % KT, c = e / KT;
            float val = 0.f;
            if (c < dk && j0 + kj < p.Tk) val = ldf<T>(kg + (int64_t)(j0 + kj) * p.ldk + c);
            KsT[c * KP + kj] = val;
        }
        for (int e = tid; e < KT * DKP; e += 128) {
            int c = e % DKP, kj = e / DKP;
            float val = 0.f;
            if (c < dk && j0 + kj < p.Tk) val = ldf<T>(vg + (int64_t)(j0 + kj) * p.ldv + c);
            Vs[kj * (DKP + 4) + c] = val;
        }
        const int omin = mlen + i0 - j0 - (KT - 1);
        if (rg) {
            for (int e = tid; e < OB * DKP; e += 128) {
                int o = e % OB, c = e / OB;
                int d = omin + o; d = d < 0 ? -d : d;
                if (p.clamp_len > 0) d = min(d, p.clamp_len);
                d = min(d, p.rlen - 1);
                float val = (c < dk) ? ldf<T>(rg + (int64_t)d * p.ldr + c) : 0.f;
                RsT[c * OP + o] = val;
            }
        }
        __syncthreads();

        // ---- S = Qu K^T (4 queries x 4 keys per thread) ----
        float S[4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int f = 0; f < 4; ++f) S[e][f] = 0.f;
#pragma unroll 8
        for (int c = 0; c < DKP; ++c) {
            float4 a = *reinterpret_cast<const float4*>(QuT + c * QP + ty * 4);
            float4 kb = *reinterpret_cast<const float4*>(KsT + c * KP + tx * 4);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {kb.x, kb.y, kb.z, kb.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int f = 0; f < 4; ++f) S[e][f] = fmaf(av[e], bv[f], S[e][f]);
        }
        // ---- BD band = Qv R^T (4 queries x 6 offsets per thread), gathered by (i - j) ----
        if (rg) {
            float BD[4][6];
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int f = 0; f < 6; ++f) BD[e][f] = 0.f;
#pragma unroll 4
            for (int c = 0; c < DKP; ++c) {
                float4 a = *reinterpret_cast<const float4*>(QvT + c * QP + ty * 4);
                float av[4] = {a.x, a.y, a.z, a.w};
                const float* rr = RsT + c * OP + tx * 6;
#pragma unroll
                for (int f = 0; f < 6; ++f) {
                    float rv = rr[f];
#pragma unroll
                    for (int e = 0; e < 4; ++e) BD[e][f] = fmaf(av[e], rv, BD[e][f]);
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int f = 0; f < 6; ++f) BDs[(ty * 4 + e) * BP + tx * 6 + f] = BD[e][f];
            __syncthreads();
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int f = 0; f < 4; ++f)
                    S[e][f] += BDs[(ty * 4 + e) * BP + (ty * 4 + e) - (tx * 4 + f) + (KT - 1)];
        }
        // ---- scale, mask, online softmax ----
        float pexp[4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = i0 + ty * 4 + e;
            float rmax = -FLT_MAX;
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const int j = j0 + tx * 4 + f;
                bool vis = j < klen;
                if (p.causal) vis = vis && (j <= mlen + i + p.lookahead);
                if (p.chunk_c > 0) {
                    int cs = ((mlen + i) / p.chunk_c) * p.chunk_c;
                    vis = vis && (j >= cs - p.chunk_l) && (j < cs + p.chunk_c);
                }
                float s = S[e][f] * p.inv_scale;
                if (!vis) s = -FLT_MAX;
                if (j >= p.Tk) s = -INFINITY;          // beyond the tensor: not part of the softmax at all
                S[e][f] = s;
                rmax = fmaxf(rmax, s);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
            const float m_new = fmaxf(m_run[e], rmax);
            const float corr = __expf(m_run[e] - m_new);
            float rsum = 0.f;
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                float pv = (S[e][f] == -INFINITY) ? 0.f : __expf(S[e][f] - m_new);
                pexp[e][f] = pv;
                rsum += pv;
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
            l_run[e] = l_run[e] * corr + rsum;
            m_run[e] = m_new;
#pragma unroll
            for (int f = 0; f < CPT; ++f) O[e][f] *= corr;
        }
        // ---- P^T to smem, O += P V ----
#pragma unroll
        for (int f = 0; f < 4; ++f)
            *reinterpret_cast<float4*>(PsT + (tx * 4 + f) * QP + ty * 4) =
                make_float4(pexp[0][f], pexp[1][f], pexp[2][f], pexp[3][f]);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < KT; ++j) {
            float4 a = *reinterpret_cast<const float4*>(PsT + j * QP + ty * 4);
            float av[4] = {a.x, a.y, a.z, a.w};
            const float* vr = Vs + j * (DKP + 4) + tx * CPT;
#pragma unroll
            for (int f = 0; f < CPT; ++f) {
                float vv = vr[f];
#pragma unroll
                for (int e = 0; e < 4; ++e) O[e][f] = fmaf(av[e], vv, O[e][f]);
            }
        }
    }
    // ---- write out ----
    T* og = reinterpret_cast<T*>(p.out) + (int64_t)b * p.Tq * p.ldo + (int64_t)h * dk;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int i = i0 + ty * 4 + e;
        if (i < p.Tq) {
            const float inv = 1.f / l_run[e];
#pragma unroll
            for (int f = 0; f < CPT; ++f) {
                int c = tx * CPT + f;
                if (c < dk) stf<T>(og + (int64_t)i * p.ldo + c, O[e][f] * inv);
            }
        }
    }
}

template <typename T, int DKP>
nsp_status launch_attn(const AttnParams& p, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)2 * DKP * QP + (size_t)DKP * KP + (size_t)KT * (DKP + 4) +
                                         (size_t)DKP * OP + (size_t)QT * BP + (size_t)KT * QP);
    auto kern = attn_simt_kernel<T, DKP>;
    static bool attr = false;
    if (!attr) { NSP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    const int qtiles = ceil_div(p.Tq, QT);
    launch_k(kern, dim3((unsigned)(p.B * p.H * qtiles)), dim3(128), smem, st, p);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

}  // namespace
}  // namespace nsp

namespace nsp {
nsp_status attention_tc_dispatch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                 const void* r, int64_t ldr, int rlen, const int32_t* klens, void* out, int64_t ldo,
                                 int B, int H, int Tq, int Tk, int dk, int clamp_len, int causal, int lookahead,
                                 int chunk_c, int chunk_l, float* stats, cudaStream_t st);
}

using namespace nsp;

static nsp_status attention_fwd_impl(int is_bf16, const void* q, int64_t ldq, const void* k, int64_t ldk,
                                     const void* v, int64_t ldv, const void* r, int64_t ldr, int rlen,
                                     const float* u_bias, const float* v_bias, const int32_t* klens,
                                     void* out, int64_t ldo, int B, int H, int Tq, int Tk, int dk,
                                     int clamp_len, int causal, int lookahead, int chunk_c, int chunk_l,
                                     float* stats, int* stats_written, void* stream) {
    if (stats_written) *stats_written = 0;
    NSP_CHECK_ARG(q && k && v && klens && out, "attention: null pointer");
    NSP_CHECK_ARG(B > 0 && H > 0 && Tq > 0 && Tk >= Tq && dk > 0, "attention: bad shape B=%d H=%d Tq=%d Tk=%d dk=%d", B, H, Tq, Tk, dk);
    NSP_CHECK_ARG(!r || rlen > 0, "attention: rlen must be positive when r is given");
    if (dk > 128) { set_error("attention: d_k=%d unsupported (max 128)", dk); return NSP_ERR_UNSUPPORTED; }
    AttnParams p;
    p.q = q; p.k = k; p.v = v; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.r = r; p.ldr = ldr; p.rlen = rlen;
    p.u_bias = u_bias; p.v_bias = v_bias; p.klens = klens; p.out = out; p.ldo = ldo;
    p.B = B; p.H = H; p.Tq = Tq; p.Tk = Tk; p.dk = dk; p.clamp_len = clamp_len; p.causal = causal;
    p.lookahead = lookahead; p.chunk_c = chunk_c; p.chunk_l = chunk_l;
    p.inv_scale = 1.0f / sqrtf((float)dk);
    cudaStream_t st = (cudaStream_t)stream;
    if (is_bf16 && !u_bias && !v_bias) {
        // tensor-core kernel when the shape is inside its envelope (d_k = 64, clamped or no relative term)
        nsp_status s = attention_tc_dispatch(q, ldq, k, ldk, v, ldv, r, ldr, rlen, klens, out, ldo, B, H, Tq, Tk, dk,
                                             clamp_len, causal, lookahead, chunk_c, chunk_l, stats, st);
        if (s != NSP_ERR_UNSUPPORTED) { if (s == NSP_OK && stats && stats_written) *stats_written = 1; return s; }
    }
    if (is_bf16) {
        if (dk <= 16) return launch_attn<__nv_bfloat16, 16>(p, st);
        if (dk <= 64) return launch_attn<__nv_bfloat16, 64>(p, st);
        return launch_attn<__nv_bfloat16, 128>(p, st);
    }
    if (dk <= 16) return launch_attn<float, 16>(p, st);
    if (dk <= 64) return launch_attn<float, 64>(p, st);
    return launch_attn<float, 128>(p, st);
}

extern "C" nsp_status nsp_relpos_attention_fwd(int is_bf16, const void* q, int64_t ldq, const void* k, int64_t ldk,
                                               const void* v, int64_t ldv, const void* r, int64_t ldr, int rlen,
                                               const float* u_bias, const float* v_bias, const int32_t* klens,
                                               void* out, int64_t ldo, int B, int H, int Tq, int Tk, int dk,
                                               int clamp_len, int causal, int lookahead, int chunk_c, int chunk_l,
                                               void* stream) {
    return attention_fwd_impl(is_bf16, q, ldq, k, ldk, v, ldv, r, ldr, rlen, u_bias, v_bias, klens, out, ldo, B, H, Tq, Tk, dk,
                              clamp_len, causal, lookahead, chunk_c, chunk_l, nullptr, nullptr, stream);
}

extern "C" nsp_status nsp_relpos_attention_fwd_stats(int is_bf16, const void* q, int64_t ldq, const void* k, int64_t ldk,
                                                     const void* v, int64_t ldv, const void* r, int64_t ldr, int rlen,
                                                     const float* u_bias, const float* v_bias, const int32_t* klens,
                                                     void* out, int64_t ldo, int B, int H, int Tq, int Tk, int dk,
                                                     int clamp_len, int causal, int lookahead, int chunk_c, int chunk_l,
                                                     float* stats, int* stats_written_host, void* stream) {
    return attention_fwd_impl(is_bf16, q, ldq, k, ldk, v, ldv, r, ldr, rlen, u_bias, v_bias, klens, out, ldo, B, H, Tq, Tk, dk,
                              clamp_len, causal, lookahead, chunk_c, chunk_l, stats, stats_written_host, stream);
}
