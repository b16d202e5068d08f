// Weight / bias gradient of the 3x3 convolutions of the VGG-style front-end on channels-last activations.
//
// The reference obtains it from autograd over nn.Conv2d in Conv2dBlock.forward (encoders/conv.py:362-394).
//   dW[co,ci,ky,kx] += sum_{b,t,f} dz[b,t,f,co] * a[b,t+ky-1,f+kx-1,ci]      dbias[co] += sum dz[b,t,f,co]
// (the input gradient is the forward kernel applied to dz with flipped, transposed taps -- see ops.conv3x3_dgrad).
// Persistent CTAs walk 8x16-position tiles; a thread owns (position slice, ci, 4 output channels) and keeps its
// 9 x 4 partial sums in registers across all of its tiles; one atomicAdd per output and CTA at the end.
#include "common.cuh"

namespace nsp {
namespace {

template <typename T> __device__ __forceinline__ float fb_ld(const T* p);
template <> __device__ __forceinline__ float fb_ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float fb_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

struct ConvWgradParams {
    const void* a; int in_chmajor;   // activations [B,T,F,CI] (or raw features [B,T,CI,F])
    const void* dz;                  // [B,T,F,CO]
    float* dw;                       // [CO,CI,3,3]
    float* dbias;                    // [CO]
    int B, T, F, CI, CO;
};

constexpr int WTH = 8, WTW = 16;

template <typename TA, typename TZ>
__global__ void __launch_bounds__(256) conv3x3_wgrad_kernel(ConvWgradParams p) {
    pdl_entry();
    extern __shared__ float sm[];
    const int CI = p.CI, CO = p.CO;
    const int cip = CI + 1;
    float* tin = sm;                                              // [(WTH+2)*(WTW+2)][cip]
    float* tz = sm + (size_t)(WTH + 2) * (WTW + 2) * cip;         // [WTH*WTW][CO]
    const int ncg = CO / 4;
    const int nps = 256 / (CI * ncg);                             // position slices (>= 1)
    const int ci = threadIdx.x % CI;
    const int cog = (threadIdx.x / CI) % ncg;
    const int ps = threadIdx.x / (CI * ncg);
    const bool active = ps < nps;
    const int ftiles = (p.F + WTW - 1) / WTW, ttiles = (p.T + WTH - 1) / WTH;
    const int64_t ntiles = (int64_t)p.B * ttiles * ftiles;

    float acc[9][4], accb[4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) accb[j] = 0.f;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int ft = (int)(tile % ftiles), tt = (int)((tile / ftiles) % ttiles), b = (int)(tile / ((int64_t)ftiles * ttiles));
        const int t0 = tt * WTH, f0 = ft * WTW;
        const TA* ag = reinterpret_cast<const TA*>(p.a) + (int64_t)b * p.T * p.F * CI;
        const TZ* zg = reinterpret_cast<const TZ*>(p.dz) + (int64_t)b * p.T * p.F * CO;
        __syncthreads();
        for (int e = threadIdx.x; e < (WTH + 2) * (WTW + 2) * CI; e += 256) {
            const int c = e % CI, pos = e / CI;
            const int ff = pos % (WTW + 2), tr = pos / (WTW + 2);
            const int t = t0 + tr - 1, f = f0 + ff - 1;
            float v = 0.f;
            if (t >= 0 && t < p.T && f >= 0 && f < p.F)
                v = p.in_chmajor ? fb_ld<TA>(ag + ((int64_t)t * CI + c) * p.F + f) : fb_ld<TA>(ag + ((int64_t)t * p.F + f) * CI + c);
            tin[pos * cip + c] = v;
        }
        for (int e = threadIdx.x; e < WTH * WTW * CO; e += 256) {
            const int co = e % CO, pos = e / CO;
            const int t = t0 + pos / WTW, f = f0 + pos % WTW;
            tz[e] = (t < p.T && f < p.F) ? fb_ld<TZ>(zg + ((int64_t)t * p.F + f) * CO + co) : 0.f;
        }
        __syncthreads();
        if (active) {
            for (int pos = ps; pos < WTH * WTW; pos += nps) {
                const int pr = pos / WTW, pc = pos % WTW;
                const float4 z4 = *reinterpret_cast<const float4*>(tz + pos * CO + cog * 4);
                const float zv[4] = {z4.x, z4.y, z4.z, z4.w};
                if (ci == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) accb[j] += zv[j];
                }
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const float av = tin[((pr + tap / 3) * (WTW + 2) + pc + tap % 3) * cip + ci];
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[tap][j] = fmaf(av, zv[j], acc[tap][j]);
                }
            }
        }
    }
    if (active) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                atomicAdd(p.dw + ((int64_t)(cog * 4 + j) * CI + ci) * 9 + tap, acc[tap][j]);
        if (ci == 0 && p.dbias) {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(p.dbias + cog * 4 + j, accb[j]);
        }
    }
}

// C_in = 1 (the raw feature map [B,T,F]): HBM-bound on dz.  One CTA walks frame rows (b, t); the three feature rows
// t-1, t, t+1 sit in shared memory; warp w takes bins f = w, w+8, ...; lane = output channel: one coalesced 64-byte dz
// read per position, nine broadcast taps, nine FMAs.  Partials stay in registers across all rows of the CTA.
template <typename TZ>
__global__ void __launch_bounds__(256) conv3x3_wgrad_c1_kernel(const float* __restrict__ x, const TZ* __restrict__ dz,
                                                               float* __restrict__ dw, float* __restrict__ dbias,
                                                               int B, int T, int F) {
    pdl_entry();
    extern __shared__ float xs[];                 // [3][F + 2]
    __shared__ float red[10][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Fp = F + 2;
    float acc[9], accb = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.f;
    const int64_t nrows = (int64_t)B * T;
    for (int64_t row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int t = (int)(row % T);
        const int64_t b = row / T;
        __syncthreads();
        for (int e = threadIdx.x; e < 3 * Fp; e += 256) {
            const int ky = e / Fp, ff = e % Fp;
            const int tt = t + ky - 1, f = ff - 1;
            xs[e] = (tt >= 0 && tt < T && f >= 0 && f < F) ? __ldg(x + ((int64_t)b * T + tt) * F + f) : 0.f;
        }
        __syncthreads();
        const TZ* zr = dz + row * (int64_t)F * 32;
        for (int f = warp; f < F; f += 8) {
            const float zv = fb_ld<TZ>(zr + (int64_t)f * 32 + lane);
            accb += zv;
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[k] = fmaf(zv, xs[(k / 3) * Fp + f + (k % 3)], acc[k]);
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 10 * 32; e += 256) (&red[0][0])[e] = 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 9; ++k) atomicAdd(&red[k][lane], acc[k]);
    atomicAdd(&red[9][lane], accb);
    __syncthreads();
    for (int e = threadIdx.x; e < 9 * 32; e += 256) atomicAdd(dw + (int64_t)(e % 32) * 9 + e / 32, red[e / 32][e % 32]);   // dw[co][0][tap]
    if (threadIdx.x < 32 && dbias) atomicAdd(dbias + threadIdx.x, red[9][threadIdx.x]);
}

}  // namespace
}  // namespace nsp

using namespace nsp;

extern "C" nsp_status nsp_conv3x3_wgrad(int a_bf16, int dz_bf16, const void* a, int in_chmajor, const void* dz, float* dw,
                                        float* dbias, int B, int T, int F, int CI, int CO, void* stream) {
    NSP_CHECK_ARG(a && dz && dw, "conv3x3_wgrad: null pointer");
    NSP_CHECK_ARG(B > 0 && T > 0 && F > 0 && CI > 0 && CO > 0, "conv3x3_wgrad: bad shape");
    // thread = (position slice, ci, 4 output channels): any CI * CO/4 <= 256 maps (threads beyond the last whole slice idle,
    // e.g. CI = 3 delta-feature planes: 10 slices of 24 threads)
    NSP_CHECK_ARG(CO % 4 == 0 && CI * (CO / 4) <= 256, "conv3x3_wgrad: CI=%d CO=%d unsupported (CO %% 4 == 0, CI*CO/4 <= 256)", CI, CO);
    if (CI == 1 && CO == 32 && !a_bf16) {           // first layer: streaming kernel ([B,T,1,F] and [B,T,F,1] coincide)
        int grid1 = 4 * num_sms();
        if ((int64_t)B * T < grid1) grid1 = B * T;
        const size_t sm1 = sizeof(float) * 3 * (size_t)(F + 2);
        if (dz_bf16) launch_k(conv3x3_wgrad_c1_kernel<__nv_bfloat16>, dim3(grid1), dim3(256), sm1, (cudaStream_t)stream, (const float*)a, (const __nv_bfloat16*)dz, dw, dbias, B, T, F);
        else launch_k(conv3x3_wgrad_c1_kernel<float>, dim3(grid1), dim3(256), sm1, (cudaStream_t)stream, (const float*)a, (const float*)dz, dw, dbias, B, T, F);
        NSP_LAUNCH_OK();
        return NSP_OK;
    }
    ConvWgradParams p;
    p.a = a; p.in_chmajor = in_chmajor; p.dz = dz; p.dw = dw; p.dbias = dbias; p.B = B; p.T = T; p.F = F; p.CI = CI; p.CO = CO;
    const size_t smem = sizeof(float) * ((size_t)(WTH + 2) * (WTW + 2) * (CI + 1) + (size_t)WTH * WTW * CO);
    const int64_t ntiles = (int64_t)B * ceil_div(T, WTH) * ceil_div(F, WTW);
    int grid = 2 * num_sms();
    if (ntiles < grid) grid = (int)ntiles;
    cudaStream_t st = (cudaStream_t)stream;
#define NSP_WG(TA, TZ)                                                                                                    \
    do {                                                                                                                  \
        auto kern = conv3x3_wgrad_kernel<TA, TZ>;                                                                         \
        static size_t attr = 0;                                                                                           \
        if (smem > 48 * 1024 && smem > attr) { NSP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; } \
        launch_k(kern, dim3(grid), dim3(256), smem, st, p);                                                                                 \
    } while (0)
    if (a_bf16 && dz_bf16) NSP_WG(__nv_bfloat16, __nv_bfloat16);
    else if (!a_bf16 && dz_bf16) NSP_WG(float, __nv_bfloat16);
    else if (!a_bf16 && !dz_bf16) NSP_WG(float, float);
    else { set_error("conv3x3_wgrad: bf16 activations with fp32 gradients are not instantiated"); return NSP_ERR_UNSUPPORTED; }
#undef NSP_WG
    NSP_LAUNCH_OK();
    return NSP_OK;
}
