// Micro-benchmark (round 2): where does the CTC lattice sweep spend its ~1 us per time step?  Standalone (no torch):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o profiles/ubench/lattice profiles/ubench/lattice.cu
// One warp sweeps alpha over T steps of S = 32 K states (K states per lane in registers, neighbours by shuffle).
//   V0  emissions staged with cp.async into shared memory (8 steps per chunk), __expf/__logf, alpha row stored per step
//   V1  emissions prefetched into registers 8 steps ahead (one vector load per step), ex2/lg2.approx.ftz in the log2
//       domain (2 ex2 + 1 lg2 per state: the max term is exactly 1), alpha row stored per step
//   V2  V1 without the per-step global store;  V3  V1 without the MUFU math (adds only): pure load / shuffle / store skeleton
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define NEG (-1.0e30f)
__device__ __forceinline__ float lse3_ref(float a, float b, float c) {
    float m = fmaxf(a, fmaxf(b, c));
    return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// log2-domain: lse2(a,b,c) = m + lg2(2^(a-m) + 2^(b-m) + 2^(c-m))
__device__ __forceinline__ float lse3_l2(float a, float b, float c) {
    const float m = fmaxf(a, fmaxf(b, c));
    return m + lg2f(ex2f(a - m) + ex2f(b - m) + ex2f(c - m));
}
__device__ __forceinline__ void cp16(void* s, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(s)), "l"(g) : "memory");
}

template <int K> __device__ __forceinline__ void ldv(const float* p, float (&v)[K]) {
    if constexpr (K >= 4) { for (int i = 0; i < K; i += 4) { float4 t = *reinterpret_cast<const float4*>(p + i); v[i] = t.x; v[i+1] = t.y; v[i+2] = t.z; v[i+3] = t.w; } }
    else if constexpr (K == 2) { float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y; }
    else v[0] = *p;
}
template <int K> __device__ __forceinline__ void stv(float* p, const float (&v)[K]) {
    if constexpr (K >= 4) { for (int i = 0; i < K; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i+1], v[i+2], v[i+3]); }
    else if constexpr (K == 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    else *p = v[0];
}

template <int K, int VAR>
__global__ void __launch_bounds__(32) sweep(const float* __restrict__ em, float* __restrict__ out, int T, float* res) {
    constexpr int Sp = 32 * K;
    __shared__ __align__(16) float sbuf[2 * 8 * Sp];
    const int lane = threadIdx.x;
    const int s0 = lane * K;
    em += (size_t)blockIdx.x * T * Sp;
    out += (size_t)blockIdx.x * T * Sp;
    float own[K];
#pragma unroll
    for (int k = 0; k < K; ++k) own[k] = NEG;
    if constexpr (VAR == 0) {
        auto stage = [&](int c) {
            float* dst = sbuf + (c & 1) * 8 * Sp;
            for (int e = lane; e < 8 * (Sp / 4); e += 32) {
                const int tt = e / (Sp / 4), v4 = e % (Sp / 4);
                if (c * 8 + tt < T) cp16(dst + tt * Sp + v4 * 4, em + (size_t)(c * 8 + tt) * Sp + v4 * 4);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        const int nch = (T + 7) / 8;
        stage(0);
        if (nch > 1) stage(1); else asm volatile("cp.async.commit_group;" ::: "memory");
        for (int c = 0; c < nch; ++c) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
            const float* eb = sbuf + (c & 1) * 8 * Sp + s0;
            const int nst = min(8, T - c * 8);
            for (int tt = 0; tt < nst; ++tt) {
                const int i = c * 8 + tt;
                float e[K], nw[K];
                ldv<K>(eb + tt * Sp, e);
                if (i == 0) {
#pragma unroll
                    for (int k = 0; k < K; ++k) nw[k] = (s0 + k <= 1) ? e[k] : NEG;
                } else {
                    float n1 = __shfl_up_sync(0xffffffffu, own[K - 1], 1);
                    float n2 = (K >= 2) ? __shfl_up_sync(0xffffffffu, own[K >= 2 ? K - 2 : 0], 1) : __shfl_up_sync(0xffffffffu, own[0], 2);
                    if (lane == 0) { n1 = NEG; n2 = NEG; }
                    if (K == 1 && lane == 1) n2 = NEG;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const float a1 = (k >= 1) ? own[k >= 1 ? k - 1 : 0] : n1;
                        float a2 = (k >= 2) ? own[k >= 2 ? k - 2 : 0] : ((k == 1) ? n1 : n2);
                        if (!((s0 + k) & 1)) a2 = NEG;
                        nw[k] = fmaxf(lse3_ref(own[k], a1, a2) + e[k], NEG);
                    }
                }
#pragma unroll
                for (int k = 0; k < K; ++k) own[k] = nw[k];
                stv<K>(out + (size_t)i * Sp + s0, own);
            }
            __syncwarp();
            if (c + 2 < nch) stage(c + 2); else asm volatile("cp.async.commit_group;" ::: "memory");
        }
    } else {
        constexpr int PF = 8;
        float ring[PF][K];
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            if (j < T) ldv<K>(em + (size_t)j * Sp + s0, ring[j]);
            else { for (int k = 0; k < K; ++k) ring[j][k] = 0.f; }
        }
        for (int i0 = 0; i0 < T; i0 += PF) {
#pragma unroll
            for (int j = 0; j < PF; ++j) {
                const int i = i0 + j;
                if (i < T) {
                    float e[K], nw[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) e[k] = ring[j][k];
                    if (i + PF < T) ldv<K>(em + (size_t)(i + PF) * Sp + s0, ring[j]);     // refill this ring slot
                    if (i == 0) {
#pragma unroll
                        for (int k = 0; k < K; ++k) nw[k] = (s0 + k <= 1) ? e[k] : NEG;
                    } else {
                        float n1 = __shfl_up_sync(0xffffffffu, own[K - 1], 1);
                        float n2 = (K >= 2) ? __shfl_up_sync(0xffffffffu, own[K >= 2 ? K - 2 : 0], 1) : __shfl_up_sync(0xffffffffu, own[0], 2);
                        if (lane == 0) { n1 = NEG; n2 = NEG; }
                        if (K == 1 && lane == 1) n2 = NEG;
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const float a1 = (k >= 1) ? own[k >= 1 ? k - 1 : 0] : n1;
                            float a2 = (k >= 2) ? own[k >= 2 ? k - 2 : 0] : ((k == 1) ? n1 : n2);
                            if (!((s0 + k) & 1)) a2 = NEG;
                            if constexpr (VAR == 3) nw[k] = fmaxf(fmaxf(own[k], fmaxf(a1, a2)) + e[k], NEG);
                            else nw[k] = fmaxf(lse3_l2(own[k], a1, a2) + e[k], NEG);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < K; ++k) own[k] = nw[k];
                    if constexpr (VAR != 2) stv<K>(out + (size_t)i * Sp + s0, own);
                }
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) acc += own[k];
    if (acc == 123.456f) res[0] = acc;      // keep the sweep alive
    if (lane == 31) res[1 + blockIdx.x] = own[K - 1];
}

template <int K, int VAR>
void run(int T, int nblk, const float* em, float* out, float* res) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) sweep<K, VAR><<<nblk, 32>>>(em, out, T, res);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) sweep<K, VAR><<<nblk, 32>>>(em, out, T, res);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    float h[2]; cudaMemcpy(h, res, sizeof(h), cudaMemcpyDeviceToHost);
    printf("K=%d var=%d T=%d blocks=%d: %.2f us per sweep, %.3f us per step   (last alpha %.4f)  %s\n", K, VAR, T, nblk, ms * 100.f,
           ms * 100.f / T, h[1], cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int T = 250, NB = 32, SpMax = 32 * 16;
    float *em, *out, *res;
    cudaMalloc(&em, (size_t)NB * T * SpMax * 4); cudaMalloc(&out, (size_t)NB * T * SpMax * 4); cudaMalloc(&res, 4096);
    float* h = (float*)malloc((size_t)NB * T * SpMax * 4);
    srand(1);
    for (size_t i = 0; i < (size_t)NB * T * SpMax; ++i) h[i] = -1.f - 8.f * (rand() / (float)RAND_MAX);
    cudaMemcpy(em, h, (size_t)NB * T * SpMax * 4, cudaMemcpyHostToDevice);
    for (int nb : {1, 32}) {
        run<1, 0>(T, nb, em, out, res); run<2, 0>(T, nb, em, out, res); run<4, 0>(T, nb, em, out, res); run<8, 0>(T, nb, em, out, res);
        run<1, 1>(T, nb, em, out, res); run<2, 1>(T, nb, em, out, res); run<4, 1>(T, nb, em, out, res); run<8, 1>(T, nb, em, out, res);
        run<4, 2>(T, nb, em, out, res); run<4, 3>(T, nb, em, out, res); run<8, 3>(T, nb, em, out, res);
    }
    return 0;
}
