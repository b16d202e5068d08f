// Shared device/host helpers for the neural_sp_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nsp_b200.h"

namespace nsp {

// ---- error plumbing (C-ABI returns nsp_status; message kept per thread) ----
void set_error(const char* fmt, ...);

#define NSP_CHECK_ARG(cond, ...)                         \
    do {                                                 \
        if (!(cond)) {                                   \
            ::nsp::set_error(__VA_ARGS__);               \
            return NSP_ERR_INVALID;                      \
        }                                                \
    } while (0)

#define NSP_CUDA_OK(expr)                                                              \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            ::nsp::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__,            \
                             cudaGetErrorName(_e), cudaGetErrorString(_e));            \
            return NSP_ERR_CUDA;                                                       \
        }                                                                              \
    } while (0)

#define NSP_LAUNCH_OK() NSP_CUDA_OK(cudaGetLastError())

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int num_sms();   // cached device query (148 on B200)
bool pdl_enabled();   // programmatic dependent launch for the kernels that opt in (NSP_PDL=0 turns it off)

// ---- programmatic dependent launch (PDL) ----
// A kernel launched through launch_k() carries cudaLaunchAttributeProgrammaticStreamSerialization: the next such kernel in
// the stream may be scheduled while this one is still running, and its CTAs wait in pdl_entry() until this grid has
// completed and flushed its memory.  What overlaps is the launch latency, the CTA rasterisation and the prologue
// (barrier init, TMEM allocation, tensormap prefetch) of kernel N+1 with the tail of kernel N -- the 3-5 us per launch that
// dominate the T' = 125 layers (profiles/README.md, round 2).  Contract for every kernel launched this way: pdl_entry()
// (or pdl_wait()) before the FIRST global-memory access that touches data another kernel writes or reads.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Same, for a thread-block cluster of `cluster` CTAs along x (grid.x must be a multiple of it).
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kc(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster,
                                    Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- device helpers ----
// thread-block cluster plumbing (distributed shared memory)
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(const float* own_smem_ptr, uint32_t rank) {     // the same variable in CTA `rank`
    uint32_t a = (uint32_t)__cvta_generic_to_shared(own_smem_ptr), ra;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Both at once, for kernels without a prologue worth overlapping: let the next kernel start its own prologue, then wait for
// the previous grid.  (Both are no-ops when the kernel was launched without the attribute.)
__device__ __forceinline__ void pdl_entry() { pdl_launch_dependents(); pdl_wait(); }
#define NSP_NEG_BIG (-1.0e30f)   // finite stand-in for log(0): avoids inf-inf NaNs in log-sum-exp sweeps

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide reductions through a 32-float smem scratch; every thread gets the result.
template <int NT>
__device__ __forceinline__ float block_max(float v, float* scratch) {
    constexpr int NW = NT / 32;
    v = warp_max(v);
    if constexpr (NW == 1) return v;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = (threadIdx.x & 31) < NW ? scratch[threadIdx.x & 31] : -INFINITY;
    return warp_max(r);
}
template <int NT>
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    constexpr int NW = NT / 32;
    v = warp_sum(v);
    if constexpr (NW == 1) return v;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = (threadIdx.x & 31) < NW ? scratch[threadIdx.x & 31] : 0.f;
    return warp_sum(r);
}

// log(exp(a)+exp(b)+exp(c)) with the NSP_NEG_BIG convention (never produces NaN).
__device__ __forceinline__ float lse3(float a, float b, float c) {
    float m = fmaxf(a, fmaxf(b, c));
    float s = __expf(a - m) + __expf(b - m) + __expf(c - m);
    return m + __logf(s);
}
__device__ __forceinline__ float lse2(float a, float b) {
    float m = fmaxf(a, b);
    return m + __logf(__expf(a - m) + __expf(b - m));
}

// streaming 128-bit accesses (read-once inputs / write-once outputs bypass L1 allocation)
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace nsp
