// Small streaming kernels: tf32 hi/lo split (parity mode operands), dtype casts.
#include "common.cuh"

namespace nsp {
namespace {

__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                                         float* __restrict__ lo, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (; i < n; i += stride) {
        float v = x[i];
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
        float h = __uint_as_float(t);
        hi[i] = h;
        lo[i] = v - h;
    }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (; i < n; i += stride) y[i] = __float2bfloat16_rn(x[i]);
}

}  // namespace
}  // namespace nsp

using namespace nsp;

static unsigned ew_grid(int64_t n) {
    int64_t b = ceil_div64(n, 256);
    int64_t cap = (int64_t)num_sms() * 16;
    return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

extern "C" nsp_status nsp_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream) {
    NSP_CHECK_ARG(x && hi && lo && n >= 0, "split_tf32: bad arguments");
    if (n == 0) return NSP_OK;
    split_tf32_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, hi, lo, n);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream) {
    NSP_CHECK_ARG(x && y && n >= 0, "cast_f32_to_bf16: bad arguments");
    if (n == 0) return NSP_OK;
    cast_bf16_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)y, n);
    NSP_LAUNCH_OK();
    return NSP_OK;
}
