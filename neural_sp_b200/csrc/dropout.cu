// Dropout for the training path (reference: nn.Dropout in every residual branch, conformer_block.py:133-180,
// positionwise_feed_forward.py:83, positional_embedding.py:139, ctc.py:87).
//
// No mask tensor exists: whether element i of call site `stream` is kept is a pure function
//     keep(i) = philox4x32_10(counter = (i / 4, stream, offset), key = seed)[i % 4]  >=  p * 2^32
// so the backward pass regenerates exactly the forward's mask from (seed, offset, stream, i) instead of re-reading 1-2 bytes
// per element, and a CUDA-graph replay sees a fresh mask as soon as the device-resident `offset` has been advanced
// (nsp_rng_advance is one captured 1-thread kernel per step).  One Philox call serves 4 consecutive elements = one
// 16-byte fp32 / 8-byte bf16 access per thread and iteration: HBM-bound streaming kernels.
//   nsp_dropout:      y = keep ? x * scale / (1 - p) : 0             (fp32 or bf16 in / out, in place allowed)
//   nsp_dropout_add:  out = res + (keep ? t * alpha / (1 - p) : 0)   (residual-branch output; fp32 residual stream)
#include "common.cuh"

namespace nsp {
namespace {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0; k.y += W1;
    }
    return c;
}

struct Rng {
    uint2 key;          // seed
    uint32_t stream;    // call-site id
    uint32_t off_lo, off_hi;
    uint32_t thresh;    // drop when r < thresh
};

__device__ __forceinline__ Rng load_rng(const unsigned long long* state, uint32_t stream, float p) {
    Rng g;
    const unsigned long long seed = state[0], off = state[1];
    g.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    g.stream = stream;
    g.off_lo = (uint32_t)off; g.off_hi = (uint32_t)(off >> 32);
    const double t = (double)p * 4294967296.0;
    g.thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
    return g;
}

// random words of element group g (elements 4g .. 4g+3)
__device__ __forceinline__ uint4 group_bits(const Rng& g, int64_t grp) {
    return philox4x32_10(make_uint4((uint32_t)grp, (uint32_t)((uint64_t)grp >> 32) ^ g.off_hi, g.stream, g.off_lo), g.key);
}

template <typename T> __device__ __forceinline__ float ldv(const T* p);
template <> __device__ __forceinline__ float ldv<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldv<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stv(T* p, float v);
template <> __device__ __forceinline__ void stv<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stv<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) dropout_kernel(const TI* x, TO* y, int64_t n, float p, float scale,
                                                      const unsigned long long* __restrict__ state, uint32_t stream) {
    pdl_entry();
    const Rng g = load_rng(state, stream, p);
    const float s = scale / (1.f - p);
    const int64_t ngrp = (n + 3) >> 2;
    for (int64_t grp = (int64_t)blockIdx.x * 256 + threadIdx.x; grp < ngrp; grp += (int64_t)gridDim.x * 256) {
        const uint4 r = group_bits(g, grp);
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
        const int64_t i0 = grp << 2;
#pragma unroll
        for (int l = 0; l < 4; ++l)
            if (i0 + l < n) stv<TO>(y + i0 + l, rr[l] >= g.thresh ? ldv<TI>(x + i0 + l) * s : 0.f);
    }
}

// 4 elements per thread as one vector access (n % 4 == 0, aligned pointers)
__global__ void __launch_bounds__(256) dropout_f32_vec_kernel(const float4* x, float4* y, int64_t ngrp,
                                                              float p, float scale, const unsigned long long* __restrict__ state,
                                                              uint32_t stream) {
    pdl_entry();
    const Rng g = load_rng(state, stream, p);
    const float s = scale / (1.f - p);
    for (int64_t grp = (int64_t)blockIdx.x * 256 + threadIdx.x; grp < ngrp; grp += (int64_t)gridDim.x * 256) {
        const uint4 r = group_bits(g, grp);
        const float4 v = x[grp];
        y[grp] = make_float4(r.x >= g.thresh ? v.x * s : 0.f, r.y >= g.thresh ? v.y * s : 0.f,
                             r.z >= g.thresh ? v.z * s : 0.f, r.w >= g.thresh ? v.w * s : 0.f);
    }
}

__global__ void __launch_bounds__(256) dropout_bf16_vec_kernel(const uint2* x, uint2* y, int64_t ngrp,
                                                               float p, float scale, const unsigned long long* __restrict__ state,
                                                               uint32_t stream) {
    pdl_entry();
    const Rng g = load_rng(state, stream, p);
    const float s = scale / (1.f - p);
    for (int64_t grp = (int64_t)blockIdx.x * 256 + threadIdx.x; grp < ngrp; grp += (int64_t)gridDim.x * 256) {
        const uint4 r = group_bits(g, grp);
        uint2 v = x[grp];
        const float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v.y));
        __nv_bfloat162 o0 = __floats2bfloat162_rn(r.x >= g.thresh ? a.x * s : 0.f, r.y >= g.thresh ? a.y * s : 0.f);
        __nv_bfloat162 o1 = __floats2bfloat162_rn(r.z >= g.thresh ? b.x * s : 0.f, r.w >= g.thresh ? b.y * s : 0.f);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&o0); o.y = *reinterpret_cast<uint32_t*>(&o1);
        y[grp] = o;
    }
}

template <typename TT>
__global__ void __launch_bounds__(256) dropout_add_kernel(const TT* __restrict__ t, const float* res,
                                                          float* out, int64_t n, float p, float alpha,
                                                          const unsigned long long* __restrict__ state, uint32_t stream) {
    pdl_entry();
    const Rng g = load_rng(state, stream, p);
    const float s = alpha / (1.f - p);
    const int64_t ngrp = (n + 3) >> 2;
    for (int64_t grp = (int64_t)blockIdx.x * 256 + threadIdx.x; grp < ngrp; grp += (int64_t)gridDim.x * 256) {
        const uint4 r = group_bits(g, grp);
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
        const int64_t i0 = grp << 2;
#pragma unroll
        for (int l = 0; l < 4; ++l)
            if (i0 + l < n) out[i0 + l] = res[i0 + l] + (rr[l] >= g.thresh ? ldv<TT>(t + i0 + l) * s : 0.f);
    }
}

__global__ void rng_advance_kernel(unsigned long long* state) {
    pdl_entry(); state[1] += 1ull; }

unsigned dr_grid(int64_t ngrp) {
    int64_t b = ceil_div64(ngrp, 256), cap = (int64_t)num_sms() * 8;
    return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace
}  // namespace nsp

using namespace nsp;

extern "C" nsp_status nsp_dropout(int in_bf16, int out_bf16, const void* x, void* y, int64_t n, float p, float scale,
                                  const uint64_t* rng_state, uint32_t stream_id, void* stream) {
    NSP_CHECK_ARG(x && y && rng_state && n >= 0 && p >= 0.f && p < 1.f, "dropout: bad arguments (p=%f)", (double)p);
    if (n == 0) return NSP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long* state = reinterpret_cast<const unsigned long long*>(rng_state);
    const int64_t ngrp = (n + 3) / 4;
    const bool vec = (n % 4 == 0) && ((((uintptr_t)x | (uintptr_t)y) & 15) == 0);
    if (!in_bf16 && !out_bf16 && vec)
        launch_k(dropout_f32_vec_kernel, dim3(dr_grid(ngrp)), dim3(256), 0, st, (const float4*)x, (float4*)y, ngrp, p, scale, state, stream_id);
    else if (in_bf16 && out_bf16 && vec)
        launch_k(dropout_bf16_vec_kernel, dim3(dr_grid(ngrp)), dim3(256), 0, st, (const uint2*)x, (uint2*)y, ngrp, p, scale, state, stream_id);
    else if (!in_bf16 && !out_bf16)
        launch_k(dropout_kernel<float, float>, dim3(dr_grid(ngrp)), dim3(256), 0, st, (const float*)x, (float*)y, n, p, scale, state, stream_id);
    else if (!in_bf16 && out_bf16)
        launch_k(dropout_kernel<float, __nv_bfloat16>, dim3(dr_grid(ngrp)), dim3(256), 0, st, (const float*)x, (__nv_bfloat16*)y, n, p, scale, state, stream_id);
    else if (in_bf16 && !out_bf16)
        launch_k(dropout_kernel<__nv_bfloat16, float>, dim3(dr_grid(ngrp)), dim3(256), 0, st, (const __nv_bfloat16*)x, (float*)y, n, p, scale, state, stream_id);
    else
        launch_k(dropout_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(dr_grid(ngrp)), dim3(256), 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, p, scale, state, stream_id);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_dropout_add(int t_bf16, const void* t, const float* res, float* out, int64_t n, float p, float alpha,
                                      const uint64_t* rng_state, uint32_t stream_id, void* stream) {
    NSP_CHECK_ARG(t && res && out && rng_state && n >= 0 && p >= 0.f && p < 1.f, "dropout_add: bad arguments (p=%f)", (double)p);
    if (n == 0) return NSP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long* state = reinterpret_cast<const unsigned long long*>(rng_state);
    const int64_t ngrp = (n + 3) / 4;
    if (t_bf16) launch_k(dropout_add_kernel<__nv_bfloat16>, dim3(dr_grid(ngrp)), dim3(256), 0, st, (const __nv_bfloat16*)t, res, out, n, p, alpha, state, stream_id);
    else launch_k(dropout_add_kernel<float>, dim3(dr_grid(ngrp)), dim3(256), 0, st, (const float*)t, res, out, n, p, alpha, state, stream_id);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_rng_advance(uint64_t* rng_state, void* stream) {
    NSP_CHECK_ARG(rng_state, "rng_advance: null state");
    launch_k(rng_advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, reinterpret_cast<unsigned long long*>(rng_state));
    NSP_LAUNCH_OK();
    return NSP_OK;
}
