// Backward of the fused Conformer convolution core  y = Swish(LayerNorm(depthwise_conv1d_k(x) + bias))   (HBM-bound).
//
// The reference obtains it from autograd over ConformerConvBlock.forward modules/conformer_convolution.py:113-124.
// Two kernels:
//   K1 (same tiling as the forward kernel): recomputes z = dwconv(x)+bias and the LayerNorm statistics in registers,
//      takes dy through Swish' and the LayerNorm backward, writes dz and adds d(norm weight / bias).
//   K2 (CTA = 32 frames x 128 channels): dx = correlation of dz with the taps, d(taps), d(conv bias).
// Only the LayerNorm variant is differentiable here (the LibriSpeech recipes' choice); BatchNorm / GroupNorm training
// is rejected by the host wrapper.
#include "common.cuh"
#include "conv_stream.h"
#include <stdlib.h>

namespace nsp {
namespace {

constexpr int RT = 4;

template <typename T> __device__ __forceinline__ float cb_ld(const T* p);
template <> __device__ __forceinline__ float cb_ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float cb_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void cb_st(T* p, float v);
template <> __device__ __forceinline__ void cb_st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void cb_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

struct ConvBwdParams {
    const void* x; int64_t ldx;      // forward input (GLU output) [B*T, d]
    const float* w;                  // [k, d] taps (transposed)
    const float* bias;               // [d]
    const float* g; const float* b;  // LayerNorm weight / bias
    const void* dy; int64_t lddy;    // gradient w.r.t. the forward output [B*T, d]
    void* dz; int64_t lddz;          // gradient w.r.t. the depthwise-conv output [B*T, d]
    void* dx; int64_t lddx;          // gradient w.r.t. x
    float* dw; float* dbias;         // [k, d], [d]  (accumulated)
    float* dg; float* db;            // [d], [d]     (accumulated)
    int B, T, d, k, left_pad;
    float eps;
};

// ---- K1 ----
template <typename T, int CPLMAX, int NW>
__global__ void __launch_bounds__(32 * NW) conv_bwd_norm_kernel(ConvBwdParams p) {
    pdl_entry();
    constexpr int NT = 32 * NW;
    constexpr int TT = NW * RT;
    extern __shared__ float sm[];
    const int d = p.d, k = p.k;
    const int rows = TT + k - 1;
    float* tile = sm;                          // [rows][d]
    float* wT = tile + (size_t)rows * d;       // [k][d]
    float* red = wT + (size_t)k * d;           // [2][d]
    const int ttiles = (p.T + TT - 1) / TT;
    const int b = blockIdx.x / ttiles, t0 = (blockIdx.x % ttiles) * TT;
    const T* xg = reinterpret_cast<const T*>(p.x) + (int64_t)b * p.T * p.ldx;
    {
        const int w_ = threadIdx.x >> 5, l_ = threadIdx.x & 31;
        const bool vec = (sizeof(T) == 2) && (d % 8 == 0) && (p.ldx % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
        for (int r = w_; r < rows; r += NW) {
            const int t = t0 + r - p.left_pad;
            const bool in = (t >= 0 && t < p.T);
            float* trow = tile + (size_t)r * d;
            if (vec) {
                for (int c8 = l_; c8 < d / 8; c8 += 32) {
                    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
                    if (in) {
                        const uint4 raw = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(xg) + (int64_t)t * p.ldx + c8 * 8);
                        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
                        const float2 a = __bfloat1622float2(h2[0]), b2 = __bfloat1622float2(h2[1]);
                        const float2 c2 = __bfloat1622float2(h2[2]), d2 = __bfloat1622float2(h2[3]);
                        lo = make_float4(a.x, a.y, b2.x, b2.y); hi = make_float4(c2.x, c2.y, d2.x, d2.y);
                    }
                    *reinterpret_cast<float4*>(trow + c8 * 8) = lo;
                    *reinterpret_cast<float4*>(trow + c8 * 8 + 4) = hi;
                }
            } else {
                for (int c = l_; c < d; c += 32) trow[c] = in ? cb_ld<T>(xg + (int64_t)t * p.ldx + c) : 0.f;
            }
        }
        for (int e = threadIdx.x; e < k * d; e += NT) wT[e] = __ldg(p.w + e);
        for (int e = threadIdx.x; e < 2 * d; e += NT) red[e] = 0.f;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tl = warp * RT;
    const int cpl = (d + 31) / 32;
    float y[CPLMAX][RT];
#pragma unroll
    for (int i = 0; i < CPLMAX; ++i) {
        const int c = lane + 32 * i;
        if (i < cpl && c < d) {
            float acc[RT];
            const float bs = __ldg(p.bias + c);
#pragma unroll
            for (int r = 0; r < RT; ++r) acc[r] = bs;
            for (int j = 0; j < k; ++j) {
                const float wv = wT[j * d + c];
#pragma unroll
                for (int r = 0; r < RT; ++r) acc[r] = fmaf(wv, tile[(tl + r + j) * d + c], acc[r]);
            }
#pragma unroll
            for (int r = 0; r < RT; ++r) y[i][r] = acc[r];
        } else {
#pragma unroll
            for (int r = 0; r < RT; ++r) y[i][r] = 0.f;
        }
    }
    float mean[RT], rstd[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < CPLMAX; ++i) s += y[i][r];
        mean[r] = warp_sum(s) / (float)d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < CPLMAX; ++i) {
            const int c = lane + 32 * i;
            if (i < cpl && c < d) { const float dd = y[i][r] - mean[r]; q += dd * dd; }
        }
        rstd[r] = rsqrtf(warp_sum(q) / (float)d + p.eps);
    }
    // ---- Swish' and LayerNorm backward (two passes over the register tile; da is recomputed in the second) ----
    const T* dyg = reinterpret_cast<const T*>(p.dy) + (int64_t)b * p.T * p.lddy;
    float s1[RT], s2[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) { s1[r] = 0.f; s2[r] = 0.f; }
#pragma unroll
    for (int i = 0; i < CPLMAX; ++i) {
        const int c = lane + 32 * i;
        const bool act = (i < cpl && c < d);
        const float gm = act ? __ldg(p.g + c) : 0.f, bt = act ? __ldg(p.b + c) : 0.f;
        float dgs = 0.f, dbs = 0.f;
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const int t = t0 + tl + r;
            const float n = act ? (y[i][r] - mean[r]) * rstd[r] : 0.f;
            const float a = n * gm + bt;
            const float sg = 1.f / (1.f + __expf(-a));
            float dyv = 0.f;
            if (act && t < p.T) dyv = cb_ld<T>(dyg + (int64_t)t * p.lddy + c);
            const float da = dyv * sg * (1.f + a * (1.f - sg));
            dgs += da * n;
            dbs += da;
            const float dn = da * gm;
            s1[r] += dn;
            s2[r] += dn * n;
        }
        if (act) { atomicAdd(red + c, dgs); atomicAdd(red + d + c, dbs); }
    }
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        s1[r] = warp_sum(s1[r]) / (float)d;
        s2[r] = warp_sum(s2[r]) / (float)d;
    }
    T* dzg = reinterpret_cast<T*>(p.dz) + (int64_t)b * p.T * p.lddz;
#pragma unroll
    for (int i = 0; i < CPLMAX; ++i) {
        const int c = lane + 32 * i;
        if (i < cpl && c < d) {
            const float gm = __ldg(p.g + c), bt = __ldg(p.b + c);
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                const int t = t0 + tl + r;
                if (t < p.T) {
                    const float n = (y[i][r] - mean[r]) * rstd[r];
                    const float a = n * gm + bt;
                    const float sg = 1.f / (1.f + __expf(-a));
                    const float dyv = cb_ld<T>(dyg + (int64_t)t * p.lddy + c);
                    const float dn = dyv * sg * (1.f + a * (1.f - sg)) * gm;
                    cb_st<T>(dzg + (int64_t)t * p.lddz + c, rstd[r] * (dn - s1[r] - n * s2[r]));
                }
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += NT) {
        if (p.dg) atomicAdd(p.dg + c, red[c]);
        if (p.db) atomicAdd(p.db + c, red[d + c]);
    }
}

// ---- K2 ----
constexpr int K2_NW = 8;
constexpr int K2_TT = K2_NW * RT;     // 32 frames
constexpr int K2_CH = 128;            // channels per CTA

template <typename T>
__global__ void __launch_bounds__(32 * K2_NW) conv_bwd_dw_kernel(ConvBwdParams p) {
    pdl_entry();
    extern __shared__ float sm[];
    const int d = p.d, k = p.k;
    const int rows = K2_TT + k - 1;
    float* xwin = sm;                                   // [rows][128]   x frames  t0 - left_pad ...
    float* zwin = xwin + (size_t)rows * K2_CH;          // [rows][128]   dz frames t0 - (k-1-left_pad) ...
    float* wch = zwin + (size_t)rows * K2_CH;           // [k][128]
    float* dwacc = wch + (size_t)k * K2_CH;             // [k][128]
    float* dbacc = dwacc + (size_t)k * K2_CH;           // [128]
    const int cchunks = (d + K2_CH - 1) / K2_CH;
    const int ttiles = (p.T + K2_TT - 1) / K2_TT;
    const int cc = blockIdx.x % cchunks;
    const int tt = (blockIdx.x / cchunks) % ttiles;
    const int b = blockIdx.x / (cchunks * ttiles);
    const int t0 = tt * K2_TT, c0 = cc * K2_CH;
    const int rpad = k - 1 - p.left_pad;
    const T* xg = reinterpret_cast<const T*>(p.x) + (int64_t)b * p.T * p.ldx;
    const T* zg = reinterpret_cast<const T*>(p.dz) + (int64_t)b * p.T * p.lddz;
    const bool vec2 = (sizeof(T) == 2) && (d % 8 == 0) && (p.ldx % 8 == 0) && (p.lddz % 8 == 0) &&
                      (((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.dz)) & 15) == 0);
    if (vec2) {
        for (int e = threadIdx.x; e < rows * (K2_CH / 8); e += 32 * K2_NW) {
            const int c8 = e % (K2_CH / 8), r = e / (K2_CH / 8);
            const int tx = t0 + r - p.left_pad, tz = t0 + r - rpad;
            const bool cok = c0 + c8 * 8 < d;
            float xv[8], zv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { xv[k] = 0.f; zv[k] = 0.f; }
            if (cok && tx >= 0 && tx < p.T) {
                const uint4 raw = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(xg) + (int64_t)tx * p.ldx + c0 + c8 * 8);
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h2[k]); xv[2 * k] = f.x; xv[2 * k + 1] = f.y; }
            }
            if (cok && tz >= 0 && tz < p.T) {
                const uint4 raw = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(zg) + (int64_t)tz * p.lddz + c0 + c8 * 8);
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h2[k]); zv[2 * k] = f.x; zv[2 * k + 1] = f.y; }
            }
            float* xd = xwin + (size_t)r * K2_CH + c8 * 8;
            float* zd = zwin + (size_t)r * K2_CH + c8 * 8;
            *reinterpret_cast<float4*>(xd) = make_float4(xv[0], xv[1], xv[2], xv[3]);
            *reinterpret_cast<float4*>(xd + 4) = make_float4(xv[4], xv[5], xv[6], xv[7]);
            *reinterpret_cast<float4*>(zd) = make_float4(zv[0], zv[1], zv[2], zv[3]);
            *reinterpret_cast<float4*>(zd + 4) = make_float4(zv[4], zv[5], zv[6], zv[7]);
        }
    } else {
        for (int e = threadIdx.x; e < rows * K2_CH; e += 32 * K2_NW) {
            const int c = e % K2_CH, r = e / K2_CH;
            const int tx = t0 + r - p.left_pad, tz = t0 + r - rpad;
            const bool cok = c0 + c < d;
            xwin[e] = (cok && tx >= 0 && tx < p.T) ? cb_ld<T>(xg + (int64_t)tx * p.ldx + c0 + c) : 0.f;
            zwin[e] = (cok && tz >= 0 && tz < p.T) ? cb_ld<T>(zg + (int64_t)tz * p.lddz + c0 + c) : 0.f;
        }
    }
    for (int e = threadIdx.x; e < k * K2_CH; e += 32 * K2_NW) {
        const int c = e % K2_CH;
        wch[e] = (c0 + c < d) ? __ldg(p.w + (int64_t)(e / K2_CH) * d + c0 + c) : 0.f;
        dwacc[e] = 0.f;
    }
    if (threadIdx.x < K2_CH) dbacc[threadIdx.x] = 0.f;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tl = warp * RT;
    T* dxg = reinterpret_cast<T*>(p.dx) + (int64_t)b * p.T * p.lddx;
#pragma unroll
    for (int i = 0; i < K2_CH / 32; ++i) {
        const int c = lane + 32 * i;
        float acc[RT], own[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) { acc[r] = 0.f; own[r] = zwin[(tl + r + rpad) * K2_CH + c]; }
        for (int j = 0; j < k; ++j) {
            const float wv = wch[j * K2_CH + c];
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                acc[r] = fmaf(wv, zwin[(tl + r + (k - 1 - j)) * K2_CH + c], acc[r]);
                s = fmaf(own[r], xwin[(tl + r + j) * K2_CH + c], s);
            }
            atomicAdd(dwacc + j * K2_CH + c, s);
        }
        atomicAdd(dbacc + c, own[0] + own[1] + own[2] + own[3]);
        if (c0 + c < d) {
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                const int t = t0 + tl + r;
                if (t < p.T) cb_st<T>(dxg + (int64_t)t * p.lddx + c0 + c, acc[r]);
            }
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < k * K2_CH; e += 32 * K2_NW) {
        const int c = e % K2_CH;
        if (c0 + c < d && p.dw) atomicAdd(p.dw + (int64_t)(e / K2_CH) * d + c0 + c, dwacc[e]);
    }
    if (threadIdx.x < K2_CH && c0 + threadIdx.x < d && p.dbias) atomicAdd(p.dbias + c0 + threadIdx.x, dbacc[threadIdx.x]);
}

template <typename T, int CPLMAX>
nsp_status launch_conv_bwd(const ConvBwdParams& p, cudaStream_t st) {
    int nw = 16;
    auto smem_of = [&](int n) { return sizeof(float) * ((size_t)(n * RT + p.k - 1) * p.d + (size_t)p.k * p.d + 2 * (size_t)p.d); };
    size_t smem = smem_of(nw);
    if (smem > 220 * 1024) { nw = 8; smem = smem_of(nw); }
    if (smem > 220 * 1024) { set_error("conformer_conv_bwd: d=%d k=%d needs %zu B smem", p.d, p.k, smem); return NSP_ERR_UNSUPPORTED; }
    const unsigned grid = (unsigned)(p.B * ceil_div(p.T, nw * RT));
    if (nw == 16) {
        auto kern = conv_bwd_norm_kernel<T, CPLMAX, 16>;
        static size_t attr = 0;
        if (smem > attr) { NSP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
        launch_k(kern, dim3(grid), dim3(512), smem, st, p);
    } else {
        auto kern = conv_bwd_norm_kernel<T, CPLMAX, 8>;
        static size_t attr = 0;
        if (smem > attr) { NSP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
        launch_k(kern, dim3(grid), dim3(256), smem, st, p);
    }
    NSP_LAUNCH_OK();
    const size_t smem2 = sizeof(float) * ((size_t)2 * (K2_TT + p.k - 1) * K2_CH + (size_t)2 * p.k * K2_CH + K2_CH);
    if (smem2 > 220 * 1024) { set_error("conformer_conv_bwd: k=%d needs %zu B smem", p.k, smem2); return NSP_ERR_UNSUPPORTED; }
    auto k2 = conv_bwd_dw_kernel<T>;
    static size_t attr2 = 0;
    if (smem2 > attr2) { NSP_CUDA_OK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); attr2 = smem2; }
    const unsigned grid2 = (unsigned)(p.B * ceil_div(p.T, K2_TT) * ceil_div(p.d, K2_CH));
    launch_k(k2, dim3(grid2), dim3(32 * K2_NW), smem2, st, p);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

}  // namespace
}  // namespace nsp

using namespace nsp;

extern "C" nsp_status nsp_conformer_conv_bwd(int is_bf16, const void* x, int64_t ldx, const float* w, const float* bias,
                                             int norm_mode, const float* norm_w, const float* norm_b, float eps,
                                             const void* dy, int64_t lddy, void* dz_ws, int64_t lddz, void* dx, int64_t lddx,
                                             float* dw, float* dbias, float* dnorm_w, float* dnorm_b,
                                             int B, int T, int d, int k, int causal, void* workspace, size_t workspace_bytes,
                                             void* stream) {
    NSP_CHECK_ARG(x && w && bias && norm_w && norm_b && dy && dz_ws && dx, "conformer_conv_bwd: null pointer");
    NSP_CHECK_ARG(B > 0 && T > 0 && d > 0 && k >= 1 && (k % 2 == 1), "conformer_conv_bwd: bad shape B=%d T=%d d=%d k=%d", B, T, d, k);
    if (norm_mode != 0) { set_error("conformer_conv_bwd: only the LayerNorm variant is differentiable on this path"); return NSP_ERR_UNSUPPORTED; }
    if (d > 1024) { set_error("conformer_conv_bwd: d=%d unsupported (max 1024)", d); return NSP_ERR_UNSUPPORTED; }
    ConvBwdParams p;
    p.x = x; p.ldx = ldx; p.w = w; p.bias = bias; p.g = norm_w; p.b = norm_b; p.dy = dy; p.lddy = lddy;
    p.dz = dz_ws; p.lddz = lddz; p.dx = dx; p.lddx = lddx; p.dw = dw; p.dbias = dbias; p.dg = dnorm_w; p.db = dnorm_b;
    p.B = B; p.T = T; p.d = d; p.k = k; p.left_pad = causal ? (k - 1) : (k - 1) / 2; p.eps = eps;
    cudaStream_t st = (cudaStream_t)stream;
    static const bool legacy = [] { const char* e = getenv("NSP_CONV_PATH"); return e && !strcmp(e, "legacy"); }();
    if (!legacy) {
        const nsp_status s = conv_stream_bwd(is_bf16, x, ldx, w, bias, norm_w, norm_b, eps, dy, lddy, dz_ws, lddz, dx, lddx, dw,
                                             dbias, dnorm_w, dnorm_b, B, T, d, k, causal, workspace, workspace_bytes, st);
        if (s != NSP_ERR_UNSUPPORTED) return s;
    }
    if (is_bf16) {
        if (d <= 256) return launch_conv_bwd<__nv_bfloat16, 8>(p, st);
        if (d <= 512) return launch_conv_bwd<__nv_bfloat16, 16>(p, st);
        return launch_conv_bwd<__nv_bfloat16, 32>(p, st);
    }
    if (d <= 256) return launch_conv_bwd<float, 8>(p, st);
    if (d <= 512) return launch_conv_bwd<float, 16>(p, st);
    return launch_conv_bwd<float, 32>(p, st);
}

extern "C" size_t nsp_conformer_conv_bwd_workspace_bytes(int B, int T, int d, int k) {
    return conv_stream_bwd_workspace_bytes(B, T, d, k);
}

// Depthwise part alone (K2): dx, d(taps), d(conv bias) from dz = gradient w.r.t. the depthwise-conv output.  Used by the
// BatchNorm training path (conformer_conv_bn.cu), whose normalisation backward needs batch-wide reductions between K1 and K2.
extern "C" nsp_status nsp_dwconv_bwd(int is_bf16, const void* x, int64_t ldx, const float* w, const void* dz, int64_t lddz,
                                     void* dx, int64_t lddx, float* dw, float* dbias, int B, int T, int d, int k, int causal,
                                     void* stream) {
    NSP_CHECK_ARG(x && w && dz && dx, "dwconv_bwd: null pointer");
    NSP_CHECK_ARG(B > 0 && T > 0 && d > 0 && k >= 1 && (k % 2 == 1), "dwconv_bwd: bad shape B=%d T=%d d=%d k=%d", B, T, d, k);
    ConvBwdParams p;
    p.x = x; p.ldx = ldx; p.w = w; p.bias = nullptr; p.g = nullptr; p.b = nullptr; p.dy = nullptr; p.lddy = 0;
    p.dz = const_cast<void*>(dz); p.lddz = lddz; p.dx = dx; p.lddx = lddx; p.dw = dw; p.dbias = dbias; p.dg = nullptr; p.db = nullptr;
    p.B = B; p.T = T; p.d = d; p.k = k; p.left_pad = causal ? (k - 1) : (k - 1) / 2; p.eps = 0.f;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem2 = sizeof(float) * ((size_t)2 * (K2_TT + k - 1) * K2_CH + (size_t)2 * k * K2_CH + K2_CH);
    if (smem2 > 220 * 1024) { set_error("dwconv_bwd: k=%d needs %zu B smem", k, smem2); return NSP_ERR_UNSUPPORTED; }
    const unsigned grid2 = (unsigned)(B * ceil_div(T, K2_TT) * ceil_div(d, K2_CH));
    if (is_bf16) {
        auto k2 = conv_bwd_dw_kernel<__nv_bfloat16>;
        NSP_CUDA_OK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        launch_k(k2, dim3(grid2), dim3(32 * K2_NW), smem2, st, p);
    } else {
        auto k2 = conv_bwd_dw_kernel<float>;
        NSP_CUDA_OK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        launch_k(k2, dim3(grid2), dim3(32 * K2_NW), smem2, st, p);
    }
    NSP_LAUNCH_OK();
    return NSP_OK;
}
