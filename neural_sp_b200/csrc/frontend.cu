// VGG-style convolutional front-end pieces on channels-last activations [B, T, F, C]:
//   conv3x3 (pad 1, stride 1) + bias + ReLU   and   max-pool (ceil_mode) with optional layout change.
//
// Replaces  Conv2dBlock.forward  encoders/conv.py:347-396  (conv1 -> relu -> conv2 -> relu -> pool) and the
// view/transposes of ConvEncoder.forward conv.py:181-189.  The reference keeps [B, C, T, F]; here channels are
// innermost so that one (t, f) position is one contiguous C-vector (coalesced 128-bit accesses) and the final
// flatten to [B, T', C*F'] (index c*F' + f, conv.py:189) is folded into the pool kernel's output indexing.
// Padded frames are not masked.
//
// This is the exact-fp32 CUDA-core path (parity mode and C_in = 1 first layer); the 32->32 layers run on
// tcgen05 as an implicit GEMM in conv_tc.cu when bf16 is selected.
#include "common.cuh"

namespace nsp {
namespace {

template <typename T> __device__ __forceinline__ float fe_ld(const T* p);
template <> __device__ __forceinline__ float fe_ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float fe_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void fe_st(T* p, float v);
template <> __device__ __forceinline__ void fe_st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void fe_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

struct Conv3Params {
    const void* x;       // input activations
    int in_chmajor;      // 1: x is [B, T, CI, F] (the reference's raw feature view, conv.py:183); 0: [B, T, F, CI]
    const float* w;      // [CO, CI, 3, 3] (nn.Conv2d weight)
    const float* bias;   // [CO]
    void* y;             // [B, T, F, CO]
    int B, T, F, CI, CO;
    int relu;
};

constexpr int TW = 16;   // tile width (frequency bins)

// 256 threads; thread = (position group of 4 consecutive bins) x (4 output channels)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) conv3x3_kernel(Conv3Params p) {
    pdl_entry();
    extern __shared__ float sm[];
    const int CI = p.CI, CO = p.CO;
    const int ncg = CO / 4;                 // channel groups
    const int npg = 256 / ncg;              // position groups per CTA
    const int TH = (npg * 4) / TW;          // tile height (frames)
    const int cip = CI + 1;                 // padded channel stride (bank spread)
    float* tin = sm;                                        // [(TH+2)*(TW+2)][cip]
    float* tw = sm + (size_t)(TH + 2) * (TW + 2) * cip;     // [9*CI][CO]
    const int ftiles = (p.F + TW - 1) / TW, ttiles = (p.T + TH - 1) / TH;
    const int ft = blockIdx.x % ftiles, tt = (blockIdx.x / ftiles) % ttiles, b = blockIdx.x / (ftiles * ttiles);
    const int t0 = tt * TH, f0 = ft * TW;
    const TI* xg = reinterpret_cast<const TI*>(p.x) + (int64_t)b * p.T * p.F * CI;

    for (int e = threadIdx.x; e < (TH + 2) * (TW + 2) * CI; e += 256) {
        int ci = e % CI, pos = e / CI;
        int ff = pos % (TW + 2), tr = pos / (TW + 2);
        int t = t0 + tr - 1, f = f0 + ff - 1;
        float v = 0.f;
        if (t >= 0 && t < p.T && f >= 0 && f < p.F)
            v = p.in_chmajor ? fe_ld<TI>(xg + ((int64_t)t * CI + ci) * p.F + f) : fe_ld<TI>(xg + ((int64_t)t * p.F + f) * CI + ci);
        tin[pos * cip + ci] = v;
    }
    for (int e = threadIdx.x; e < 9 * CI * CO; e += 256) {
        int co = e % CO, kk = e / CO;            // kk = (ky*3+kx)*CI + ci
        int ci = kk % CI, tap = kk / CI;
        tw[e] = __ldg(p.w + ((int64_t)co * CI + ci) * 9 + tap);
    }
    __syncthreads();

    const int cg = threadIdx.x % ncg, pg = threadIdx.x / ncg;
    const int pr = (pg * 4) / TW, pc = (pg * 4) % TW;      // tile-local frame row / first bin
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap % 3;
        const float* ip = tin + ((pr + ky) * (TW + 2) + pc + kx) * cip;
        const float* wp = tw + (size_t)tap * CI * CO + cg * 4;
        for (int ci = 0; ci < CI; ++ci) {
            const float4 wv = *reinterpret_cast<const float4*>(wp + (size_t)ci * CO);
            const float a0 = ip[ci], a1 = ip[cip + ci], a2 = ip[2 * cip + ci], a3 = ip[3 * cip + ci];
            acc[0][0] = fmaf(a0, wv.x, acc[0][0]); acc[0][1] = fmaf(a0, wv.y, acc[0][1]);
            acc[0][2] = fmaf(a0, wv.z, acc[0][2]); acc[0][3] = fmaf(a0, wv.w, acc[0][3]);
            acc[1][0] = fmaf(a1, wv.x, acc[1][0]); acc[1][1] = fmaf(a1, wv.y, acc[1][1]);
            acc[1][2] = fmaf(a1, wv.z, acc[1][2]); acc[1][3] = fmaf(a1, wv.w, acc[1][3]);
            acc[2][0] = fmaf(a2, wv.x, acc[2][0]); acc[2][1] = fmaf(a2, wv.y, acc[2][1]);
            acc[2][2] = fmaf(a2, wv.z, acc[2][2]); acc[2][3] = fmaf(a2, wv.w, acc[2][3]);
            acc[3][0] = fmaf(a3, wv.x, acc[3][0]); acc[3][1] = fmaf(a3, wv.y, acc[3][1]);
            acc[3][2] = fmaf(a3, wv.z, acc[3][2]); acc[3][3] = fmaf(a3, wv.w, acc[3][3]);
        }
    }
    const int t = t0 + pr;
    if (t < p.T) {
        TO* yg = reinterpret_cast<TO*>(p.y) + (((int64_t)b * p.T + t) * p.F) * CO;
        const float4 bv = *reinterpret_cast<const float4*>(p.bias + cg * 4);
        const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = f0 + pc + i;
            if (f < p.F) {
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v = acc[i][j] + bb[j];
                    o[j] = p.relu ? fmaxf(v, 0.f) : v;
                }
                TO* dst = yg + (int64_t)f * CO + cg * 4;
                if constexpr (sizeof(TO) == 2) {          // 4 bf16 = one 8-byte store
                    __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b2 = __floats2bfloat162_rn(o[2], o[3]);
                    uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&a); pk.y = *reinterpret_cast<uint32_t*>(&b2);
                    *reinterpret_cast<uint2*>(dst) = pk;
                } else {
                    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
    }
}

struct PoolParams {
    const void* x;   // [B, T, F, C]
    void* y;         // out_chmajor ? [B, To, C*Fo] (index c*Fo + f, conv.py:189) : [B, To, Fo, C]
    int B, T, F, C, pt, pf, To, Fo, Fo_keep;
    int out_chmajor;
};

// max-pool kernel=stride=(pt,pf), padding 0, ceil_mode=True (conv.py:330-337): edge windows are clipped
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) maxpool_kernel(PoolParams p) {
    pdl_entry();
    const int64_t n = (int64_t)p.B * p.To * p.Fo_keep * p.C;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
        int c = (int)(e % p.C);
        int64_t r = e / p.C;
        int fo = (int)(r % p.Fo_keep); r /= p.Fo_keep;
        int to = (int)(r % p.To);
        int b = (int)(r / p.To);
        float m = -INFINITY;
        for (int dt = 0; dt < p.pt; ++dt) {
            int t = to * p.pt + dt;
            if (t >= p.T) break;
            for (int df = 0; df < p.pf; ++df) {
                int f = fo * p.pf + df;
                if (f >= p.F) break;
                m = fmaxf(m, fe_ld<TI>(reinterpret_cast<const TI*>(p.x) + (((int64_t)b * p.T + t) * p.F + f) * p.C + c));
            }
        }
        TO* yg = reinterpret_cast<TO*>(p.y);
        if (p.out_chmajor) fe_st<TO>(yg + ((int64_t)b * p.To + to) * ((int64_t)p.C * p.Fo_keep) + (int64_t)c * p.Fo_keep + fo, m);
        else fe_st<TO>(yg + (((int64_t)b * p.To + to) * p.Fo_keep + fo) * p.C + c, m);
    }
}

// 1-D pooling over time on [B, T, D] rows, kernel = stride = factor, ceil_mode (windows clipped at T):
// mode 0 max (MaxPoolSubsampler subsampling.py:175-209), 1 mean over the in-range frames (MeanPoolSubsampler :212-246),
// 2 first frame (DropSubsampler :97-126), 3 sum (AddSubsampler :129-172)
template <typename T>
__global__ void __launch_bounds__(256) pool_time_kernel(const T* x, T* y, int B, int Tin, int Tout, int D, int factor, int mode) {
    pdl_entry();
    const int64_t n = (int64_t)B * Tout * D;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
        int c = (int)(e % D);
        int64_t r = e / D;
        int to = (int)(r % Tout), b = (int)(r / Tout);
        float m = (mode == 0) ? -INFINITY : 0.f;
        int cnt = 0;
        for (int dt = 0; dt < factor; ++dt) {
            int t = to * factor + dt;
            if (t >= Tin) break;
            float v = fe_ld<T>(x + ((int64_t)b * Tin + t) * D + c);
            if (mode == 0) m = fmaxf(m, v);
            else if (mode == 2) { if (dt == 0) m = v; }
            else m += v;
            ++cnt;
        }
        if (mode == 1) m /= (float)cnt;
        fe_st<T>(y + e, m);
    }
}

}  // namespace
}  // namespace nsp

using namespace nsp;

static unsigned fe_grid(int64_t n) {
    int64_t b = ceil_div64(n, 256), cap = (int64_t)num_sms() * 16;
    return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

extern "C" nsp_status nsp_conv3x3_relu_fwd(int in_bf16, int out_bf16, const void* x, int in_chmajor, const float* w,
                                           const float* bias, void* y, int B, int T, int F, int CI, int CO, int relu,
                                           void* stream) {
    NSP_CHECK_ARG(x && w && bias && y, "conv3x3: null pointer");
    NSP_CHECK_ARG(B > 0 && T > 0 && F > 0 && CI > 0 && CO > 0, "conv3x3: bad shape");
    if (CO % 4 != 0 || CO > 256 || (256 % (CO / 4)) != 0 || ((256 / (CO / 4)) * 4) % TW != 0) {
        set_error("conv3x3: C_out=%d unsupported (need 4 | C_out, C_out/4 | 256, C_out <= 64)", CO);
        return NSP_ERR_UNSUPPORTED;
    }
    Conv3Params p;
    p.x = x; p.in_chmajor = in_chmajor; p.w = w; p.bias = bias; p.y = y; p.B = B; p.T = T; p.F = F; p.CI = CI; p.CO = CO; p.relu = relu;
    const int npg = 256 / (CO / 4), TH = npg * 4 / TW;
    const size_t smem = sizeof(float) * ((size_t)(TH + 2) * (TW + 2) * (CI + 1) + (size_t)9 * CI * CO);
    if (smem > 220 * 1024) { set_error("conv3x3: CI=%d CO=%d needs %zu B smem", CI, CO, smem); return NSP_ERR_UNSUPPORTED; }
    const unsigned grid = (unsigned)(B * ceil_div(T, TH) * ceil_div(F, TW));
    cudaStream_t st = (cudaStream_t)stream;
#define NSP_C3(TI, TO)                                                                                   \
    do {                                                                                                 \
        auto kern = conv3x3_kernel<TI, TO>;                                                              \
        NSP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        launch_k(kern, dim3(grid), dim3(256), smem, st, p);                                                                \
    } while (0)
    if (!in_bf16 && !out_bf16) NSP_C3(float, float);
    else if (!in_bf16 && out_bf16) NSP_C3(float, __nv_bfloat16);
    else if (in_bf16 && out_bf16) NSP_C3(__nv_bfloat16, __nv_bfloat16);
    else NSP_C3(__nv_bfloat16, float);
#undef NSP_C3
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_maxpool2d_fwd(int in_bf16, int out_bf16, const void* x, void* y, int B, int T, int F, int C,
                                        int pool_t, int pool_f, int f_keep, int out_chmajor, void* stream) {
    NSP_CHECK_ARG(x && y && B > 0 && T > 0 && F > 0 && C > 0 && pool_t > 0 && pool_f > 0, "maxpool2d: bad arguments");
    PoolParams p;
    p.x = x; p.y = y; p.B = B; p.T = T; p.F = F; p.C = C; p.pt = pool_t; p.pf = pool_f;
    p.To = ceil_div(T, pool_t); p.Fo = ceil_div(F, pool_f);
    p.Fo_keep = f_keep > 0 ? f_keep : p.Fo;
    NSP_CHECK_ARG(p.Fo_keep <= p.Fo, "maxpool2d: f_keep=%d > %d", f_keep, p.Fo);
    p.out_chmajor = out_chmajor;
    const int64_t n = (int64_t)B * p.To * p.Fo_keep * C;
    cudaStream_t st = (cudaStream_t)stream;
    if (!in_bf16 && !out_bf16) launch_k(maxpool_kernel<float, float>, dim3(fe_grid(n)), dim3(256), 0, st, p);
    else if (in_bf16 && out_bf16) launch_k(maxpool_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(fe_grid(n)), dim3(256), 0, st, p);
    else if (!in_bf16) launch_k(maxpool_kernel<float, __nv_bfloat16>, dim3(fe_grid(n)), dim3(256), 0, st, p);
    else launch_k(maxpool_kernel<__nv_bfloat16, float>, dim3(fe_grid(n)), dim3(256), 0, st, p);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_pool_time_fwd(int is_bf16, const void* x, void* y, int B, int T, int D, int factor, int mode, void* stream) {
    NSP_CHECK_ARG(x && y && B > 0 && T > 0 && D > 0 && factor > 0 && mode >= 0 && mode <= 3, "pool_time: bad arguments");
    const int To = ceil_div(T, factor);
    const int64_t n = (int64_t)B * To * D;
    cudaStream_t st = (cudaStream_t)stream;
    if (is_bf16) launch_k(pool_time_kernel<__nv_bfloat16>, dim3(fe_grid(n)), dim3(256), 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, B, T, To, D, factor, mode);
    else launch_k(pool_time_kernel<float>, dim3(fe_grid(n)), dim3(256), 0, st, (const float*)x, (float*)y, B, T, To, D, factor, mode);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_maxpool_time_fwd(int is_bf16, const void* x, void* y, int B, int T, int D, int factor, void* stream) {
    return nsp_pool_time_fwd(is_bf16, x, y, B, T, D, factor, 0, stream);
}
