// Conformer convolution module, middle part, fused (HBM-bound):
//     y = Swish( Norm( depthwise_conv1d_k(x) + bias ) )       x, y: [B, T, d]  (time-major rows, d contiguous)
//
// Replaces  ConformerConvBlock.forward  modules/conformer_convolution.py:113-124
//   (depthwise nn.Conv1d(groups=d, padding=(k-1)/2 | causal k-1 with right trim :114-115), then
//    nn.LayerNorm(d, eps=1e-12) | BatchNorm1d (eval: running stats) | GroupNorm(d/2 groups), then Swish)
// and removes the three transposes/contiguous copies around it (:109,:116,:128): the pointwise
// convolutions run as row-major GEMMs (gemm_tcgen05.cu, GLU fused in the first one's epilogue).
// Padding semantics: zeros outside [0, T) of the PADDED batch tensor; frames beyond an utterance's
// own length are NOT masked (reference behaviour).
//
// One CTA = 16 warps = 64 consecutive frames of one utterance; the (32 + k - 1) x d input window and the
// k x d taps sit in shared memory; a warp owns 4 frames, a lane owns channels lane, lane+32, ...;
// the per-frame normalisation statistics are warp-shuffle reductions over registers.
#include "common.cuh"
#include "conv_stream.h"
#include <stdlib.h>

namespace nsp {
namespace {

constexpr int RT = 4;          // frames per warp
// warps per CTA: 16 (64 frames per CTA, one CTA per SM, more latency hiding) or 8 when the window would not fit in smem

struct ConvParams {
    const void* x; int64_t ldx;      // [B*T, d]
    const float* w;                  // [k, d] depthwise taps, TRANSPOSED from nn.Conv1d weight [d,1,k]
    const float* bias;               // [d]
    const float* g; const float* b;  // norm weight / bias [d]
    const float* rmean; const float* rvar;   // BatchNorm running stats (mode 1)
    void* y; int64_t ldy;
    int B, T, d, k, left_pad, mode;  // mode 0 LayerNorm, 1 BatchNorm(eval), 2 GroupNorm(2 channels per group)
    float eps;
};

template <typename T> __device__ __forceinline__ float cv_ld(const T* p);
template <> __device__ __forceinline__ float cv_ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float cv_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void cv_st(T* p, float v);
template <> __device__ __forceinline__ void cv_st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void cv_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T, int CPLMAX, int K, int NW>   // K == 0: runtime kernel size
__global__ void __launch_bounds__(32 * NW) conformer_conv_kernel(ConvParams p) {
    pdl_entry();
    constexpr int NT = 32 * NW;
    constexpr int TT = NW * RT;    // frames per CTA
    extern __shared__ float sm[];
    const int d = p.d, k = (K > 0) ? K : p.k;
    const int rows = TT + k - 1;
    float* tile = sm;                 // [rows][d]
    float* wT = sm + (size_t)rows * d;  // [k][d]
    const int ttiles = (p.T + TT - 1) / TT;
    const int b = blockIdx.x / ttiles, t0 = (blockIdx.x % ttiles) * TT;
    const T* xg = reinterpret_cast<const T*>(p.x) + (int64_t)b * p.T * p.ldx;

    {
        // stage the (TT + k - 1) x d input window (fp32 in smem) and the k x d taps; 128-bit loads when d % 8 == 0
        const int w_ = threadIdx.x >> 5, l_ = threadIdx.x & 31;
        const bool vec = (d % 8 == 0) && (p.ldx % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
        for (int r = w_; r < rows; r += NW) {
            const int t = t0 + r - p.left_pad;
            const bool in = (t >= 0 && t < p.T);
            float* trow = tile + (size_t)r * d;
            if (vec) {
                if constexpr (sizeof(T) == 2) {
                    for (int c8 = l_; c8 < d / 8; c8 += 32) {
                        float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
                        if (in) {
                            uint4 raw = *reinterpret_cast<const uint4*>(xg + (int64_t)t * p.ldx + c8 * 8);
                            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
                            float2 a = __bfloat1622float2(h2[0]), b2 = __bfloat1622float2(h2[1]);
                            float2 c2 = __bfloat1622float2(h2[2]), d2 = __bfloat1622float2(h2[3]);
                            lo = make_float4(a.x, a.y, b2.x, b2.y); hi = make_float4(c2.x, c2.y, d2.x, d2.y);
                        }
                        *reinterpret_cast<float4*>(trow + c8 * 8) = lo;
                        *reinterpret_cast<float4*>(trow + c8 * 8 + 4) = hi;
                    }
                } else {
                    for (int c4 = l_; c4 < d / 4; c4 += 32) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (in) v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(xg) + (int64_t)t * p.ldx + c4 * 4);
                        *reinterpret_cast<float4*>(trow + c4 * 4) = v;
                    }
                }
            } else {
                for (int c = l_; c < d; c += 32) trow[c] = in ? cv_ld<T>(xg + (int64_t)t * p.ldx + c) : 0.f;
            }
        }
        // taps arrive pre-transposed [k][d] (host prepares them once): straight coalesced copy
        for (int e = threadIdx.x; e < k * d; e += NT) wT[e] = __ldg(p.w + e);
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tl = warp * RT;                       // first local frame of this warp
    float y[CPLMAX][RT];
    const int cpl = (d + 31) / 32;
#pragma unroll
    for (int i = 0; i < CPLMAX; ++i) {
        const int c = lane + 32 * i;
        if (i < cpl && c < d) {
            float acc[RT];
            const float bs = __ldg(p.bias + c);
#pragma unroll
            for (int r = 0; r < RT; ++r) acc[r] = bs;
            if constexpr (K > 0) {
                float xw[RT + K - 1];
#pragma unroll
                for (int j = 0; j < RT + K - 1; ++j) xw[j] = tile[(tl + j) * d + c];
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const float wv = wT[j * d + c];
#pragma unroll
                    for (int r = 0; r < RT; ++r) acc[r] = fmaf(wv, xw[r + j], acc[r]);
                }
            } else {
                for (int j = 0; j < k; ++j) {
                    const float wv = wT[j * d + c];
#pragma unroll
                    for (int r = 0; r < RT; ++r) acc[r] = fmaf(wv, tile[(tl + r + j) * d + c], acc[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < RT; ++r) y[i][r] = acc[r];
        } else {
#pragma unroll
            for (int r = 0; r < RT; ++r) y[i][r] = 0.f;
        }
    }

    // ---- normalisation ----
    float mean[RT], rstd[RT];
    if (p.mode == 0) {
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < CPLMAX; ++i) s += y[i][r];      // inactive channels hold 0
            mean[r] = warp_sum(s) / (float)d;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < CPLMAX; ++i) {
                const int c = lane + 32 * i;
                if (i < cpl && c < d) { float dd = y[i][r] - mean[r]; q += dd * dd; }
            }
            rstd[r] = rsqrtf(warp_sum(q) / (float)d + p.eps);
        }
    }
    T* yg = reinterpret_cast<T*>(p.y) + (int64_t)b * p.T * p.ldy;
#pragma unroll
    for (int i = 0; i < CPLMAX; ++i) {
        const int c = lane + 32 * i;
        const bool act = (i < cpl && c < d);
        const float gm = act ? __ldg(p.g + c) : 0.f, bt = act ? __ldg(p.b + c) : 0.f;
        float mu_c = 0.f, rs_c = 1.f;
        if (p.mode == 1 && act) { mu_c = __ldg(p.rmean + c); rs_c = rsqrtf(__ldg(p.rvar + c) + p.eps); }
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const int t = t0 + tl + r;
            float v;
            if (p.mode == 0) v = (y[i][r] - mean[r]) * rstd[r];
            else if (p.mode == 1) v = (y[i][r] - mu_c) * rs_c;
            else {
                // GroupNorm with 2 channels per group: partner channel c^1 lives in the neighbouring lane
                // (shuffle executed by the whole warp: mode is CTA-uniform)
                float other = __shfl_xor_sync(0xffffffffu, y[i][r], 1);
                float mu = 0.5f * (y[i][r] + other);
                float d0 = y[i][r] - mu, d1 = other - mu;
                v = d0 * rsqrtf(0.5f * (d0 * d0 + d1 * d1) + p.eps);
            }
            v = v * gm + bt;
            v = __fdividef(v, 1.f + __expf(-v));     // Swish
            if (act && t < p.T) cv_st<T>(yg + (int64_t)t * p.ldy + c, v);
        }
    }
}

template <typename T, int CPLMAX>
nsp_status launch_conv(const ConvParams& p, cudaStream_t st) {
    int nw = 16;
    size_t smem = sizeof(float) * ((size_t)(nw * RT + p.k - 1) * p.d + (size_t)p.k * p.d);
    if (smem > 220 * 1024) { nw = 8; smem = sizeof(float) * ((size_t)(nw * RT + p.k - 1) * p.d + (size_t)p.k * p.d); }
    if (smem > 220 * 1024) { set_error("conformer_conv: d=%d k=%d needs %zu B smem", p.d, p.k, smem); return NSP_ERR_UNSUPPORTED; }
    const unsigned grid = (unsigned)(p.B * ceil_div(p.T, nw * RT));
#define NSP_CONV(KK)                                                                                          \
    do {                                                                                                      \
        if (nw == 16) {                                                                                       \
            auto kern = conformer_conv_kernel<T, CPLMAX, KK, 16>;                                             \
            static size_t attr_smem = 0;                                                                      \
            if (smem > attr_smem) {                                                                           \
                NSP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                attr_smem = smem;                                                                             \
            }                                                                                                 \
            launch_k(kern, dim3(grid), dim3(512), smem, st, p);                                                                 \
        } else {                                                                                              \
            auto kern = conformer_conv_kernel<T, CPLMAX, KK, 8>;                                              \
            static size_t attr_smem = 0;                                                                      \
            if (smem > attr_smem) {                                                                           \
                NSP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                attr_smem = smem;                                                                             \
            }                                                                                                 \
            launch_k(kern, dim3(grid), dim3(256), smem, st, p);                                                                 \
        }                                                                                                     \
    } while (0)
    if (p.k == 15) NSP_CONV(15); else if (p.k == 31) NSP_CONV(31); else if (p.k == 7) NSP_CONV(7); else NSP_CONV(0);
#undef NSP_CONV
    NSP_LAUNCH_OK();
    return NSP_OK;
}

}  // namespace
}  // namespace nsp

using namespace nsp;

extern "C" nsp_status nsp_conformer_conv_fwd(int is_bf16, const void* x, int64_t ldx, const float* w, const float* bias,
                                             int norm_mode, const float* norm_w, const float* norm_b,
                                             const float* run_mean, const float* run_var, float eps,
                                             void* y, int64_t ldy, int B, int T, int d, int k, int causal, void* stream) {
    NSP_CHECK_ARG(x && w && bias && norm_w && norm_b && y, "conformer_conv: null pointer");
    NSP_CHECK_ARG(B > 0 && T > 0 && d > 0 && k >= 1 && (k % 2 == 1), "conformer_conv: bad shape B=%d T=%d d=%d k=%d", B, T, d, k);
    NSP_CHECK_ARG(norm_mode >= 0 && norm_mode <= 2, "conformer_conv: norm_mode=%d", norm_mode);
    NSP_CHECK_ARG(norm_mode != 1 || (run_mean && run_var), "conformer_conv: BatchNorm needs running stats");
    NSP_CHECK_ARG(norm_mode != 2 || d % 2 == 0, "conformer_conv: GroupNorm(2 per group) needs even d");
    ConvParams p;
    p.x = x; p.ldx = ldx; p.w = w; p.bias = bias; p.g = norm_w; p.b = norm_b; p.rmean = run_mean; p.rvar = run_var;
    p.y = y; p.ldy = ldy; p.B = B; p.T = T; p.d = d; p.k = k; p.left_pad = causal ? (k - 1) : (k - 1) / 2;
    p.mode = norm_mode; p.eps = eps;
    cudaStream_t st = (cudaStream_t)stream;
    static const bool legacy = [] { const char* e = getenv("NSP_CONV_PATH"); return e && !strcmp(e, "legacy"); }();
    if (norm_mode == 0 && !legacy) {        // LayerNorm variant: streaming kernel when the shape is covered
        const nsp_status s = conv_stream_fwd(is_bf16, x, ldx, w, bias, norm_w, norm_b, eps, y, ldy, B, T, d, k, causal, st);
        if (s != NSP_ERR_UNSUPPORTED) return s;
    }
    if (d > 1024) { set_error("conformer_conv: d=%d unsupported (max 1024)", d); return NSP_ERR_UNSUPPORTED; }
    if (is_bf16) {
        if (d <= 256) return launch_conv<__nv_bfloat16, 8>(p, st);
        if (d <= 512) return launch_conv<__nv_bfloat16, 16>(p, st);
        return launch_conv<__nv_bfloat16, 32>(p, st);
    }
    if (d <= 256) return launch_conv<float, 8>(p, st);
    if (d <= 512) return launch_conv<float, 16>(p, st);
    return launch_conv<float, 32>(p, st);
}
