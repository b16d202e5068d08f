// Training of the Conformer convolution module with BatchNorm (the reference's DEFAULT `conformer_normalization`;
// modules/conformer_convolution.py:118-124: BatchNorm1d applied to a [B*T, d, 1] view, i.e. per-channel statistics over
// ALL B*T frames of the batch, padded ones included).  HBM-bound streaming kernels around the existing ones:
//
//   forward   z = dwconv_k(x) + bias                      nsp_dwconv_stats_fwd: writes z and accumulates sum z, sum z^2 per channel
//             (mu, var) = batch statistics of z            (d-length host-side finalisation; running statistics updated there)
//             y = Swish(g (z - mu) rsqrt(var + eps) + b)   the EXISTING fused forward kernel with (mu, var) as its statistics
//   backward  du = dy Swish'(u),  u = g zh + b,  zh = (z - mu) rstd
//             s1 = sum du, s2 = sum du zh  (per channel)   nsp_bn_swish_bwd_reduce   (d g = s2, d b = s1)
//             dz = g rstd (du - s1 / M - zh s2 / M)        nsp_bn_swish_bwd_apply
//             dx, d taps, d bias from dz                    nsp_dwconv_bwd = the depthwise backward kernel of conformer_conv_bwd.cu
#include "common.cuh"

namespace nsp {
namespace {

template <typename T> __device__ __forceinline__ float bn_ld(const T* p);
template <> __device__ __forceinline__ float bn_ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float bn_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void bn_st(T* p, float v);
template <> __device__ __forceinline__ void bn_st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void bn_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T> __device__ __forceinline__ float bn_round(float v);          // value as it will read back from a T
template <> __device__ __forceinline__ float bn_round<float>(float v) { return v; }
template <> __device__ __forceinline__ float bn_round<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

constexpr int CH = 128;     // channels per CTA (one thread each: coalesced rows)
constexpr int TT = 32;      // frames per CTA

// z[b, t, c] = bias[c] + sum_j w[j, c] x[b, t + j - left_pad, c]; stats[c] += sum z, stats[d + c] += sum z^2 over the tile.
// The statistics are taken on the value that is STORED (rounded to T), so that the backward sees the same z.
template <typename T>
__global__ void __launch_bounds__(CH) dwconv_stats_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                          const float* __restrict__ bias, T* __restrict__ z, int64_t ldz,
                                                          float* __restrict__ stats, int B, int Tn, int d, int k, int left_pad) {
    pdl_entry();
    const int cchunks = (d + CH - 1) / CH;
    const int ttiles = (Tn + TT - 1) / TT;
    const int cc = blockIdx.x % cchunks;
    const int tt = (blockIdx.x / cchunks) % ttiles;
    const int b = blockIdx.x / (cchunks * ttiles);
    const int c = cc * CH + threadIdx.x;
    if (c >= d) return;
    const T* xb = x + (int64_t)b * Tn * ldx + c;
    T* zb = z + (int64_t)b * Tn * ldz + c;
    const float bv = __ldg(bias + c);
    float s = 0.f, ss = 0.f;
    const int t1 = min(Tn, (tt + 1) * TT);
    for (int t = tt * TT; t < t1; ++t) {
        float acc = bv;
        for (int j = 0; j < k; ++j) {
            const int ti = t + j - left_pad;
            if (ti >= 0 && ti < Tn) acc = fmaf(__ldg(w + (int64_t)j * d + c), bn_ld<T>(xb + (int64_t)ti * ldx), acc);
        }
        bn_st<T>(zb + (int64_t)t * ldz, acc);
        const float zr = bn_round<T>(acc);                       // the stored (possibly bf16-rounded) value
        s += zr; ss = fmaf(zr, zr, ss);
    }
    atomicAdd(stats + c, s);
    atomicAdd(stats + d + c, ss);
}

__device__ __forceinline__ float swish_grad(float u) {
    const float sg = 1.f / (1.f + __expf(-u));
    return sg * (1.f + u * (1.f - sg));
}

// sums[c] += sum_rows du, sums[d + c] += sum_rows du * zh   over a tile of 64 rows
// SWISH = false: the upstream gradient is already the gradient w.r.t. the normalised value u (BatchNorm2d + ReLU blocks of the CNN
// front-end hand in the ReLU / pooling-masked gradient), beta is not read.
template <typename T, bool SWISH = true>
__global__ void __launch_bounds__(CH) bn_swish_bwd_reduce_kernel(const T* __restrict__ z, int64_t ldz, const T* __restrict__ dy,
                                                                 int64_t lddy, const float* __restrict__ mean,
                                                                 const float* __restrict__ var, const float* __restrict__ g,
                                                                 const float* __restrict__ bta, float eps, float* __restrict__ sums,
                                                                 int64_t M, int d) {
    pdl_entry();
    const int cchunks = (d + CH - 1) / CH;
    const int cc = blockIdx.x % cchunks;
    const int64_t r0 = (int64_t)(blockIdx.x / cchunks) * 64;
    const int c = cc * CH + threadIdx.x;
    if (c >= d) return;
    const float mu = __ldg(mean + c), rstd = rsqrtf(__ldg(var + c) + eps), gv = __ldg(g + c), bv = SWISH ? __ldg(bta + c) : 0.f;
    float s1 = 0.f, s2 = 0.f;
    const int64_t r1 = min(M, r0 + 64);
    for (int64_t r = r0; r < r1; ++r) {
        const float zh = (bn_ld<T>(z + r * ldz + c) - mu) * rstd;
        const float du = bn_ld<T>(dy + r * lddy + c) * (SWISH ? swish_grad(fmaf(gv, zh, bv)) : 1.f);
        s1 += du; s2 = fmaf(du, zh, s2);
    }
    atomicAdd(sums + c, s1);
    atomicAdd(sums + d + c, s2);
}

// dz = g rstd (du - s1 / M - zh s2 / M)
template <typename T, bool SWISH = true>
__global__ void __launch_bounds__(256) bn_swish_bwd_apply_kernel(const T* __restrict__ z, int64_t ldz, const T* __restrict__ dy,
                                                                 int64_t lddy, const float* __restrict__ mean,
                                                                 const float* __restrict__ var, const float* __restrict__ g,
                                                                 const float* __restrict__ bta, float eps,
                                                                 const float* __restrict__ sums, T* __restrict__ dz, int64_t lddz,
                                                                 int64_t M, int d) {
    pdl_entry();
    const int64_t n = M * d;
    const float invM = 1.f / (float)M;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
        const int c = (int)(e % d);
        const int64_t r = e / d;
        const float rstd = rsqrtf(__ldg(var + c) + eps), gv = __ldg(g + c);
        const float zh = (bn_ld<T>(z + r * ldz + c) - __ldg(mean + c)) * rstd;
        const float du = bn_ld<T>(dy + r * lddy + c) * (SWISH ? swish_grad(fmaf(gv, zh, __ldg(bta + c))) : 1.f);
        bn_st<T>(dz + r * lddz + c, gv * rstd * (du - __ldg(sums + c) * invM - zh * __ldg(sums + d + c) * invM));
    }
}

// GroupNorm with 2 channels per group on the [B*T, d, 1] view (conformer_convolution.py:47-48, :118-124: statistics per frame and
// per channel pair, no batch coupling): backward of Swish(g zh + b) through the pair statistics.  One thread = one channel pair,
// a CTA = 128 pairs x 64 rows; d g / d b are reduced per thread over its rows, then one atomicAdd each.
template <typename T>
__global__ void __launch_bounds__(CH) gn2_swish_bwd_kernel(const T* __restrict__ z, int64_t ldz, const T* __restrict__ dy,
                                                           int64_t lddy, const float* __restrict__ g, const float* __restrict__ bta,
                                                           float eps, T* __restrict__ dz, int64_t lddz, float* __restrict__ dgam,
                                                           float* __restrict__ dbet, int64_t M, int d) {
    pdl_entry();
    const int pairs = d / 2;
    const int pchunks = (pairs + CH - 1) / CH;
    const int pc = blockIdx.x % pchunks;
    const int64_t r0 = (int64_t)(blockIdx.x / pchunks) * 64;
    const int pr = pc * CH + threadIdx.x;
    if (pr >= pairs) return;
    const int c0 = 2 * pr, c1 = c0 + 1;
    const float g0 = __ldg(g + c0), g1 = __ldg(g + c1), b0 = __ldg(bta + c0), b1 = __ldg(bta + c1);
    float dg0 = 0.f, dg1 = 0.f, db0 = 0.f, db1 = 0.f;
    const int64_t r1 = min(M, r0 + 64);
    for (int64_t r = r0; r < r1; ++r) {
        const float a = bn_ld<T>(z + r * ldz + c0), b = bn_ld<T>(z + r * ldz + c1);
        const float m = 0.5f * (a + b);
        const float da = a - m, dbv = b - m;
        const float rstd = rsqrtf(0.5f * (da * da + dbv * dbv) + eps);
        const float zh0 = da * rstd, zh1 = dbv * rstd;
        const float du0 = bn_ld<T>(dy + r * lddy + c0) * swish_grad(fmaf(g0, zh0, b0));
        const float du1 = bn_ld<T>(dy + r * lddy + c1) * swish_grad(fmaf(g1, zh1, b1));
        dg0 = fmaf(du0, zh0, dg0); dg1 = fmaf(du1, zh1, dg1); db0 += du0; db1 += du1;
        const float e0 = du0 * g0, e1 = du1 * g1;                       // gradient w.r.t. zh
        const float me = 0.5f * (e0 + e1), mez = 0.5f * (e0 * zh0 + e1 * zh1);
        bn_st<T>(dz + r * lddz + c0, rstd * (e0 - me - zh0 * mez));
        bn_st<T>(dz + r * lddz + c1, rstd * (e1 - me - zh1 * mez));
    }
    atomicAdd(dgam + c0, dg0); atomicAdd(dgam + c1, dg1);
    atomicAdd(dbet + c0, db0); atomicAdd(dbet + c1, db1);
}

}  // namespace
}  // namespace nsp

using namespace nsp;

extern "C" nsp_status nsp_gn2_swish_bwd(int is_bf16, const void* z, int64_t ldz, const void* dy, int64_t lddy, const float* gamma,
                                        const float* beta, float eps, void* dz, int64_t lddz, float* dgamma, float* dbeta,
                                        int64_t M, int d, void* stream) {
    NSP_CHECK_ARG(z && dy && gamma && beta && dz && dgamma && dbeta && M > 0 && d > 0 && d % 2 == 0, "gn2_swish_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)(ceil_div64(M, 64) * ceil_div(d / 2, CH));
    if (is_bf16) launch_k(gn2_swish_bwd_kernel<__nv_bfloat16>, dim3(grid), dim3(CH), 0, st, (const __nv_bfloat16*)z, ldz, (const __nv_bfloat16*)dy, lddy, gamma, beta, eps, (__nv_bfloat16*)dz, lddz, dgamma, dbeta, M, d);
    else launch_k(gn2_swish_bwd_kernel<float>, dim3(grid), dim3(CH), 0, st, (const float*)z, ldz, (const float*)dy, lddy, gamma, beta, eps, (float*)dz, lddz, dgamma, dbeta, M, d);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_dwconv_stats_fwd(int is_bf16, const void* x, int64_t ldx, const float* w, const float* bias, void* z,
                                           int64_t ldz, float* stats, int B, int T, int d, int k, int causal, void* stream) {
    NSP_CHECK_ARG(x && w && bias && z && stats, "dwconv_stats_fwd: null pointer");
    NSP_CHECK_ARG(B > 0 && T > 0 && d > 0 && k >= 1 && (k % 2 == 1), "dwconv_stats_fwd: bad shape B=%d T=%d d=%d k=%d", B, T, d, k);
    cudaStream_t st = (cudaStream_t)stream;
    NSP_CUDA_OK(cudaMemsetAsync(stats, 0, 2 * (size_t)d * sizeof(float), st));
    const unsigned grid = (unsigned)((int64_t)B * ceil_div(T, TT) * ceil_div(d, CH));
    const int left_pad = causal ? (k - 1) : (k - 1) / 2;
    if (is_bf16) launch_k(dwconv_stats_kernel<__nv_bfloat16>, dim3(grid), dim3(CH), 0, st, (const __nv_bfloat16*)x, ldx, w, bias, (__nv_bfloat16*)z, ldz, stats, B, T, d, k, left_pad);
    else launch_k(dwconv_stats_kernel<float>, dim3(grid), dim3(CH), 0, st, (const float*)x, ldx, w, bias, (float*)z, ldz, stats, B, T, d, k, left_pad);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_bn_swish_bwd(int is_bf16, const void* z, int64_t ldz, const void* dy, int64_t lddy, const float* mean,
                                       const float* var, const float* gamma, const float* beta, float eps, float* sums,
                                       void* dz, int64_t lddz, int64_t M, int d, void* stream) {
    NSP_CHECK_ARG(z && dy && mean && var && gamma && beta && sums && dz && M > 0 && d > 0, "bn_swish_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    NSP_CUDA_OK(cudaMemsetAsync(sums, 0, 2 * (size_t)d * sizeof(float), st));
    const unsigned g1 = (unsigned)(ceil_div64(M, 64) * ceil_div(d, CH));
    int64_t b2 = ceil_div64(M * d, 256), cap = (int64_t)num_sms() * 16;
    const unsigned g2 = (unsigned)(b2 < cap ? b2 : cap);
    if (is_bf16) {
        launch_k(bn_swish_bwd_reduce_kernel<__nv_bfloat16>, dim3(g1), dim3(CH), 0, st, (const __nv_bfloat16*)z, ldz, (const __nv_bfloat16*)dy, lddy, mean, var, gamma, beta, eps, sums, M, d);
        launch_k(bn_swish_bwd_apply_kernel<__nv_bfloat16>, dim3(g2), dim3(256), 0, st, (const __nv_bfloat16*)z, ldz, (const __nv_bfloat16*)dy, lddy, mean, var, gamma, beta, eps, sums, (__nv_bfloat16*)dz, lddz, M, d);
    } else {
        launch_k(bn_swish_bwd_reduce_kernel<float>, dim3(g1), dim3(CH), 0, st, (const float*)z, ldz, (const float*)dy, lddy, mean, var, gamma, beta, eps, sums, M, d);
        launch_k(bn_swish_bwd_apply_kernel<float>, dim3(g2), dim3(256), 0, st, (const float*)z, ldz, (const float*)dy, lddy, mean, var, gamma, beta, eps, sums, (float*)dz, lddz, M, d);
    }
    NSP_LAUNCH_OK();
    return NSP_OK;
}

// BatchNorm backward with batch statistics and NO activation inside (du = the gradient w.r.t. the normalised value): the
// BatchNorm2d + ReLU blocks of the CNN front-end in training (encoders/conv.py:362-394, nn.BatchNorm2d over [B, C, T, F] =
// per-channel statistics over the M = B*T*F rows of the channels-last activation).  sums = (d beta, d gamma).
extern "C" nsp_status nsp_bn_bwd(int is_bf16, const void* z, int64_t ldz, const void* du, int64_t lddu, const float* mean,
                                 const float* var, const float* gamma, float eps, float* sums, void* dz, int64_t lddz,
                                 int64_t M, int d, void* stream) {
    NSP_CHECK_ARG(z && du && mean && var && gamma && sums && dz && M > 0 && d > 0, "bn_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    NSP_CUDA_OK(cudaMemsetAsync(sums, 0, 2 * (size_t)d * sizeof(float), st));
    const unsigned g1 = (unsigned)(ceil_div64(M, 64) * ceil_div(d, CH));
    int64_t b2 = ceil_div64(M * d, 256), cap = (int64_t)num_sms() * 16;
    const unsigned g2 = (unsigned)(b2 < cap ? b2 : cap);
    if (is_bf16) {
        launch_k(bn_swish_bwd_reduce_kernel<__nv_bfloat16, false>, dim3(g1), dim3(CH), 0, st, (const __nv_bfloat16*)z, ldz, (const __nv_bfloat16*)du, lddu, mean, var, gamma, nullptr, eps, sums, M, d);
        launch_k(bn_swish_bwd_apply_kernel<__nv_bfloat16, false>, dim3(g2), dim3(256), 0, st, (const __nv_bfloat16*)z, ldz, (const __nv_bfloat16*)du, lddu, mean, var, gamma, nullptr, eps, sums, (__nv_bfloat16*)dz, lddz, M, d);
    } else {
        launch_k(bn_swish_bwd_reduce_kernel<float, false>, dim3(g1), dim3(CH), 0, st, (const float*)z, ldz, (const float*)du, lddu, mean, var, gamma, nullptr, eps, sums, M, d);
        launch_k(bn_swish_bwd_apply_kernel<float, false>, dim3(g2), dim3(256), 0, st, (const float*)z, ldz, (const float*)du, lddu, mean, var, gamma, nullptr, eps, sums, (float*)dz, lddz, M, d);
    }
    NSP_LAUNCH_OK();
    return NSP_OK;
}
