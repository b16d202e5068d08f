// Streaming (HBM-bound) backward kernels of the encoder path: LayerNorm backward with the residual gradient fused in,
// activation / GLU derivatives, accumulating column sums (bias gradients), time max-pool backward, ReLU masks and the
// 2-D max-pool backward of the convolutional front-end.
//
// The reference obtains all of these from autograd over
//   nn.LayerNorm                       encoders/conformer_block.py:53-80, encoders/transformer.py:600
//   Swish / ReLU / GELU, F.glu         modules/positionwise_feed_forward.py:77-89, modules/conformer_convolution.py:110-112
//   nn.Linear / nn.Conv1d biases       (column sums of the output gradient)
//   MaxPoolSubsampler (ceil_mode)      encoders/subsampling.py:175-209
//   ReLU + nn.MaxPool2d(ceil_mode)     encoders/conv.py:362-394
#include "common.cuh"

namespace nsp {
namespace {

template <typename T> __device__ __forceinline__ float bw_ld(const T* p);
template <> __device__ __forceinline__ float bw_ld<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float bw_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void bw_st(T* p, float v);
template <> __device__ __forceinline__ void bw_st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void bw_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  y = (x*s - mean) * rstd * gamma + beta
//   dx = s * rstd * (g - mean(g) - xhat * mean(g * xhat)) + dres,  g = dy * gamma
//   dgamma += sum_rows dy * xhat ;  dbeta += sum_rows dy
// One warp per row (row in registers, statistics recomputed from x exactly like the forward kernel), rows are
// grid-strided so that each thread keeps its dgamma / dbeta partials in registers; one smem reduction over the
// 8 warps and one atomicAdd per column per CTA at the end.
// ------------------------------------------------------------------------------------------------
template <int VPT>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, int64_t lddy,
                                                            const float* __restrict__ x, int64_t ldx,
                                                            const float* __restrict__ gamma, float eps, float in_scale,
                                                            const float* __restrict__ dres, int64_t lddr,
                                                            float* __restrict__ dx, int64_t lddx,
                                                            __nv_bfloat16* __restrict__ dxb, int64_t lddxb,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            float* __restrict__ dcol, float dcol_alpha,
                                                            int M, int D) {
    pdl_entry();
    extern __shared__ float red[];     // [8][D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float gsum[VPT][4], bsum[VPT][4], gm[VPT][4], csum[VPT][4];
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        const int idx = (j * 32 + lane) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            gsum[j][k] = 0.f; bsum[j][k] = 0.f; csum[j][k] = 0.f;
            gm[j][k] = (idx + k < D) ? __ldg(gamma + idx + k) : 0.f;
        }
    }
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < M; row += (int64_t)gridDim.x * 8) {
        const float* xr = x + row * ldx;
        const float* dyr = dy + row * lddy;
        float v[VPT][4], g[VPT][4];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int idx = (j * 32 + lane) * 4;
            if (idx < D) {      // D % 4 == 0 is required by the host wrapper
                const float4 t = *reinterpret_cast<const float4*>(xr + idx);
                const float4 u = *reinterpret_cast<const float4*>(dyr + idx);
                v[j][0] = t.x * in_scale; v[j][1] = t.y * in_scale; v[j][2] = t.z * in_scale; v[j][3] = t.w * in_scale;
                g[j][0] = u.x; g[j][1] = u.y; g[j][2] = u.z; g[j][3] = u.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) { v[j][k] = 0.f; g[j][k] = 0.f; }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) s += v[j][k];
        }
        const float mean = warp_sum(s) / (float)D;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int idx = (j * 32 + lane) * 4;
            if (idx < D) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float d = v[j][k] - mean; q += d * d; }
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int idx = (j * 32 + lane) * 4;
            if (idx < D) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float xh = (v[j][k] - mean) * rstd;
                    const float dyv = g[j][k];
                    gsum[j][k] += dyv * xh;
                    bsum[j][k] += dyv;
                    const float gg = dyv * gm[j][k];
                    v[j][k] = xh;
                    g[j][k] = gg;
                    s1 += gg;
                    s2 += gg * xh;
                }
            }
        }
        const float m1 = warp_sum(s1) / (float)D, m2 = warp_sum(s2) / (float)D;
        const float sc = rstd * in_scale;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int idx = (j * 32 + lane) * 4;
            if (idx < D) {
                float o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] = sc * (g[j][k] - m1 - v[j][k] * m2);
                if (dres) {
                    const float4 r = *reinterpret_cast<const float4*>(dres + row * lddr + idx);
                    o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) csum[j][k] += o[k];
                if (dx) *reinterpret_cast<float4*>(dx + row * lddx + idx) = make_float4(o[0], o[1], o[2], o[3]);
                if (dxb) {
                    __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b = __floats2bfloat162_rn(o[2], o[3]);
                    uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&a); pk.y = *reinterpret_cast<uint32_t*>(&b);
                    *reinterpret_cast<uint2*>(dxb + row * lddxb + idx) = pk;
                }
            }
        }
    }
    // ---- CTA reduction of the parameter gradients: three rounds through one [8][D] buffer (keeps shared memory, and with
    //      it the number of resident CTAs, independent of how many sums are formed).  Tried in round 2 and REJECTED: summing the
    //      CTAs' partials over distributed shared memory in clusters of 8 (8x fewer global atomics) -- 2.27 -> 2.98 ms per
    //      training step: gang-scheduling 8 CTAs and two cluster barriers cost more than the atomics they save. ----
#pragma unroll 1
    for (int round = 0; round < 3; ++round) {
        float* out = round == 0 ? dgamma : (round == 1 ? dbeta : dcol);
        if (out == nullptr) continue;                       // uniform
        float* rg = red + (size_t)warp * D;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int idx = (j * 32 + lane) * 4;
            if (idx < D) {
#pragma unroll
                for (int k = 0; k < 4; ++k) rg[idx + k] = round == 0 ? gsum[j][k] : (round == 1 ? bsum[j][k] : csum[j][k]);
            }
        }
        __syncthreads();
        const float sc_out = round == 2 ? dcol_alpha : 1.f;
        for (int c = threadIdx.x; c < D; c += 256) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += red[(size_t)w * D + c];
            atomicAdd(out + c, sc_out * t);
        }
        __syncthreads();
    }
}

// Any D, pitch and alignment (rows the register-resident kernel does not take: D % 4 != 0 such as a 10-wide last projection,
// D > 2048): one warp per row, the row is re-read from L1 / L2 for each of the three statistics, parameter gradients go
// straight to global atomics.  Odd shapes only -- not a tuned path.
__global__ void __launch_bounds__(256) layernorm_bwd_generic_kernel(const float* __restrict__ dy, int64_t lddy,
                                                                    const float* __restrict__ x, int64_t ldx,
                                                                    const float* __restrict__ gamma, float eps, float in_scale,
                                                                    const float* __restrict__ dres, int64_t lddr,
                                                                    float* __restrict__ dx, int64_t lddx,
                                                                    __nv_bfloat16* __restrict__ dxb, int64_t lddxb,
                                                                    float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                    float* __restrict__ dcol, float dcol_alpha, int M, int D) {
    pdl_entry();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < M; row += (int64_t)gridDim.x * 8) {
        const float* xr = x + row * ldx;
        const float* dyr = dy + row * lddy;
        float s = 0.f;
        for (int c = lane; c < D; c += 32) s += xr[c] * in_scale;
        const float mean = warp_sum(s) / (float)D;
        float q = 0.f;
        for (int c = lane; c < D; c += 32) { const float d = xr[c] * in_scale - mean; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
        float s1 = 0.f, s2 = 0.f;
        for (int c = lane; c < D; c += 32) {
            const float xh = (xr[c] * in_scale - mean) * rstd;
            const float gg = dyr[c] * __ldg(gamma + c);
            s1 += gg;
            s2 = fmaf(gg, xh, s2);
        }
        const float m1 = warp_sum(s1) / (float)D, m2 = warp_sum(s2) / (float)D;
        const float sc = rstd * in_scale;
        for (int c = lane; c < D; c += 32) {
            const float xh = (xr[c] * in_scale - mean) * rstd;
            const float dyv = dyr[c];
            float o = sc * (dyv * __ldg(gamma + c) - m1 - xh * m2);
            if (dres) o += dres[row * lddr + c];
            if (dx) dx[row * lddx + c] = o;
            if (dxb) dxb[row * lddxb + c] = __float2bfloat16_rn(o);
            if (dgamma) atomicAdd(dgamma + c, dyv * xh);
            if (dbeta) atomicAdd(dbeta + c, dyv);
            if (dcol) atomicAdd(dcol + c, dcol_alpha * o);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// dz = dh * act'(z)    (act: 1 relu, 2 swish, 3 gelu(erf), 4 gelu(tanh))
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dact(int act, float z) {
    if (act == 1) return z > 0.f ? 1.f : 0.f;
    if (act == 2) { const float s = 1.f / (1.f + __expf(-z)); return s * (1.f + z * (1.f - s)); }
    if (act == 3) {
        const float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752f));
        return cdf + z * 0.3989422804014327f * __expf(-0.5f * z * z);
    }
    if (act == 4) {
        const float u = 0.79788456080286536f * (z + 0.044715f * z * z * z);
        const float t = tanhf(u);
        return 0.5f * (1.f + t) + 0.5f * z * (1.f - t * t) * 0.79788456080286536f * (1.f + 3.f * 0.044715f * z * z);
    }
    return 1.f;
}

template <typename T>
__global__ void __launch_bounds__(256) act_bwd_kernel(const T* __restrict__ dh, const T* __restrict__ z, T* __restrict__ dz,
                                                      int64_t n, int act) {
    pdl_entry();
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        bw_st<T>(dz + i, bw_ld<T>(dh + i) * dact(act, bw_ld<T>(z + i)));
}

// 16-byte vector variants for bf16 tensors (8 elements per thread and iteration; n % 8 == 0, 16-byte aligned pointers)
__device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]) {
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 t = __bfloat1622float2(h2[k]); f[2 * k] = t.x; f[2 * k + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 o;
    __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]), d = __floats2bfloat162_rn(f[6], f[7]);
    o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
    o.z = *reinterpret_cast<uint32_t*>(&c); o.w = *reinterpret_cast<uint32_t*>(&d);
    return o;
}

__global__ void __launch_bounds__(256) act_bwd_vec_kernel(const uint4* __restrict__ dh, const uint4* __restrict__ z,
                                                          uint4* __restrict__ dz, int64_t n8, int act) {
    pdl_entry();
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (int64_t)gridDim.x * 256) {
        float a[8], b[8];
        unpack8(dh[i], a); unpack8(z[i], b);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] *= dact(act, b[k]);
        dz[i] = pack8(a);
    }
}

// GLU backward, bf16, one thread = 8 consecutive channels of one row (d % 8 == 0)
__global__ void __launch_bounds__(256) glu_bwd_vec_kernel(const uint4* __restrict__ dg, const uint4* __restrict__ pre,
                                                          uint4* __restrict__ dpre, int64_t M, int d8) {
    pdl_entry();
    const int64_t n = M * d8;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int64_t r = i / d8;
        const int c = (int)(i - r * d8);
        float g[8], a[8], b[8], oa[8], ob[8];
        unpack8(dg[i], g); unpack8(pre[r * 2 * d8 + c], a); unpack8(pre[r * 2 * d8 + d8 + c], b);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float sgm = 1.f / (1.f + __expf(-b[k]));
            oa[k] = g[k] * sgm;
            ob[k] = g[k] * a[k] * sgm * (1.f - sgm);
        }
        dpre[r * 2 * d8 + c] = pack8(oa);
        dpre[r * 2 * d8 + d8 + c] = pack8(ob);
    }
}

// Row-structured variants that also accumulate dbias[c] += sum_rows(result[., c]) (the bias gradient of the layer whose
// pre-activation this is): thread = 8 fixed columns, the 8 warps of a CTA stride over the rows of its slice.
// MODE 0: dz = dh * act'(z) on [M, N];  MODE 1: GLU backward, dg [M, N/2], pre / dpre [M, N].
template <int MODE>
__global__ void __launch_bounds__(256) act_bwd_bias_kernel(const __nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ z,
                                                           __nv_bfloat16* __restrict__ dz, float* __restrict__ dbias,
                                                           int M, int N, int act) {
    pdl_entry();
    __shared__ float part[8][257];
    const int lane = threadIdx.x & 31, rs = threadIdx.x >> 5;
    const int c0 = (blockIdx.x * 32 + lane) * 8;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    if (c0 < N) {
        for (int r = blockIdx.y * 8 + rs; r < M; r += gridDim.y * 8) {
            float o[8];
            if constexpr (MODE == 0) {
                float a[8], b[8];
                unpack8(*reinterpret_cast<const uint4*>(dh + (int64_t)r * N + c0), a);
                unpack8(*reinterpret_cast<const uint4*>(z + (int64_t)r * N + c0), b);
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = a[k] * dact(act, b[k]);
            } else {
                const int d = N / 2;
                const bool gate = c0 >= d;                        // columns [d, 2d) are the gate half
                const int c = gate ? c0 - d : c0;
                float g[8], a[8], b[8];
                unpack8(*reinterpret_cast<const uint4*>(dh + (int64_t)r * d + c), g);
                unpack8(*reinterpret_cast<const uint4*>(z + (int64_t)r * N + c), a);
                unpack8(*reinterpret_cast<const uint4*>(z + (int64_t)r * N + d + c), b);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float sgm = 1.f / (1.f + __expf(-b[k]));
                    o[k] = gate ? g[k] * a[k] * sgm * (1.f - sgm) : g[k] * sgm;
                }
            }
            *reinterpret_cast<uint4*>(dz + (int64_t)r * N + c0) = pack8(o);
            // the bias gradient sums the ROUNDED values the weight-gradient GEMM will read
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += o[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) part[rs][lane * 8 + k] = acc[k];
    __syncthreads();
    const int col = blockIdx.x * 256 + threadIdx.x;
    if (col < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
        atomicAdd(dbias + col, t);
    }
}

__global__ void __launch_bounds__(256) relu_mask_vec_kernel(const uint4* __restrict__ dx, const uint4* __restrict__ a,
                                                            uint4* __restrict__ dz, int64_t n8) {
    pdl_entry();
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (int64_t)gridDim.x * 256) {
        float x[8], y[8];
        unpack8(dx[i], x); unpack8(a[i], y);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = y[k] > 0.f ? x[k] : 0.f;
        dz[i] = pack8(x);
    }
}

// GLU backward: out = a * sigmoid(b) with pre = [a | b] of width 2*d per row.
template <typename T>
__global__ void __launch_bounds__(256) glu_bwd_kernel(const T* __restrict__ dg, const T* __restrict__ pre, T* __restrict__ dpre,
                                                      int64_t M, int d) {
    pdl_entry();
    const int64_t n = M * d;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int64_t r = i / d;
        const int c = (int)(i - r * d);
        const float a = bw_ld<T>(pre + r * 2 * d + c), b = bw_ld<T>(pre + r * 2 * d + d + c);
        const float g = bw_ld<T>(dg + i);
        const float s = 1.f / (1.f + __expf(-b));
        bw_st<T>(dpre + r * 2 * d + c, g * s);
        bw_st<T>(dpre + r * 2 * d + d + c, g * a * s * (1.f - s));
    }
}

// y[n] += alpha * sum_m x[m, n]; grid = (column blocks, row slices).  Vector variant: a thread owns VEC consecutive
// columns (16-byte loads), a warp covers 32*VEC columns of one row, the 8 warps stride over the rows of the slice.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) colsum_acc_vec_kernel(const T* __restrict__ x, int64_t ldx, float* __restrict__ y,
                                                             int M, int N, float alpha) {
    pdl_entry();
    __shared__ float part[8][32 * VEC + 1];
    const int lane = threadIdx.x & 31, rs = threadIdx.x >> 5;
    const int c0 = (blockIdx.x * 32 + lane) * VEC;
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
    if (c0 < N) {                                     // N % VEC == 0 (host)
        for (int r = blockIdx.y * 8 + rs; r < M; r += gridDim.y * 8) {
            const T* p = x + (int64_t)r * ldx + c0;
            if constexpr (sizeof(T) == 2) {
                const uint4 raw = *reinterpret_cast<const uint4*>(p);
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h2[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
            } else {
                const float4 f = *reinterpret_cast<const float4*>(p);
                acc[0] += f.x; acc[1] += f.y; acc[2] += f.z; acc[3] += f.w;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) part[rs][lane * VEC + k] = acc[k];
    __syncthreads();
    for (int c = threadIdx.x; c < 32 * VEC; c += 256) {
        const int col = blockIdx.x * 32 * VEC + c;
        if (col < N) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) t += part[i][c];
            atomicAdd(y + col, alpha * t);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) colsum_acc_kernel(const T* __restrict__ x, int64_t ldx, float* __restrict__ y,
                                                         int M, int N, float alpha) {
    pdl_entry();
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), rs = threadIdx.x >> 5;
    float acc = 0.f;
    if (c < N)
        for (int r = blockIdx.y * 8 + rs; r < M; r += gridDim.y * 8) acc += bw_ld<T>(x + (int64_t)r * ldx + c);
    part[rs][threadIdx.x & 31] = acc;
    __syncthreads();
    if (rs == 0 && c < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x & 31];
        atomicAdd(y + c, alpha * t);
    }
}

// MaxPoolSubsampler backward (kernel = stride = factor, ceil_mode): the gradient of output frame to goes to the FIRST
// frame of its window holding the maximum (torch's max_pool1d tie rule), zeros elsewhere.  x, dx fp32 [B,T,D].
__global__ void __launch_bounds__(256) maxpool_time_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               float* __restrict__ dx, int B, int T, int D, int factor) {
    pdl_entry();
    const int To = (T + factor - 1) / factor;
    const int64_t n = (int64_t)B * To * D;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
        const int c = (int)(e % D);
        const int64_t r = e / D;
        const int to = (int)(r % To), b = (int)(r / To);
        const int t0 = to * factor, t1 = min(T, t0 + factor);
        const float* xp = x + ((int64_t)b * T) * D + c;
        float best = xp[(int64_t)t0 * D];
        int arg = t0;
        for (int t = t0 + 1; t < t1; ++t) {
            const float v = xp[(int64_t)t * D];
            if (v > best) { best = v; arg = t; }
        }
        const float g = dy[e];
        for (int t = t0; t < t1; ++t) dx[((int64_t)b * T + t) * D + c] = (t == arg) ? g : 0.f;
    }
}

// Mean / drop / add time-pooling backward (kernel = stride = factor, ceil_mode; forward: pool_time_kernel in frontend.cu):
// every input frame t belongs to exactly one window to = t / factor.  mode 1 mean: dy / (frames of the window inside the
// input), 2 drop: dy to the window's first frame only, 3 add: dy to every frame.  dy fp32 [B,ceil(T/f),D], dx fp32 [B,T,D].
__global__ void __launch_bounds__(256) pool_time_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int T,
                                                            int D, int factor, int mode) {
    pdl_entry();
    const int To = (T + factor - 1) / factor;
    const int64_t n = (int64_t)B * T * D;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
        const int c = (int)(e % D);
        const int64_t r = e / D;
        const int t = (int)(r % T), b = (int)(r / T);
        const int to = t / factor;
        float g = __ldg(dy + ((int64_t)b * To + to) * D + c);
        if (mode == 1) g /= (float)(min(T, (to + 1) * factor) - to * factor);
        else if (mode == 2) g = (t == to * factor) ? g : 0.f;
        dx[e] = g;
    }
}

// dz = (a > 0) ? dx : 0   (ReLU backward through the saved post-activation a)
template <typename T>
__global__ void __launch_bounds__(256) relu_mask_kernel(const T* __restrict__ dx, const T* __restrict__ a, T* __restrict__ dz, int64_t n) {
    pdl_entry();
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        bw_st<T>(dz + i, bw_ld<T>(a + i) > 0.f ? bw_ld<T>(dx + i) : 0.f);
}

// ReLU + MaxPool2d(ceil_mode, kernel = stride = (pt, pf)) backward on channels-last [B,T,F,C]:
//   dz[b,t,f,c] = dy[b,to,fo,c] if (t,f) is the first arg-max of its window and a > 0 else 0.
// a is the saved post-ReLU activation; dy is [B,To,Fo,C] (in_chmajor=0) or the flattened [B,To,C*Fo] (index c*Fo+fo).
template <typename T, typename TG>
__global__ void __launch_bounds__(256) maxpool2d_relu_bwd_kernel(const T* __restrict__ a, const TG* __restrict__ dy, T* __restrict__ dz,
                                                                 int B, int Tn, int F, int C, int pt, int pf, int in_chmajor) {
    pdl_entry();
    const int To = (Tn + pt - 1) / pt, Fo = (F + pf - 1) / pf;
    const int64_t n = (int64_t)B * To * Fo * C;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
        const int c = (int)(e % C);
        int64_t r = e / C;
        const int fo = (int)(r % Fo); r /= Fo;
        const int to = (int)(r % To), b = (int)(r / To);
        const int t0 = to * pt, t1 = min(Tn, t0 + pt), f0 = fo * pf, f1 = min(F, f0 + pf);
        float best = -INFINITY;
        int at = t0, af = f0;
        for (int t = t0; t < t1; ++t)
            for (int f = f0; f < f1; ++f) {
                const float v = bw_ld<T>(a + (((int64_t)b * Tn + t) * F + f) * C + c);
                if (v > best) { best = v; at = t; af = f; }
            }
        const float g = in_chmajor ? bw_ld<TG>(dy + ((int64_t)b * To + to) * ((int64_t)C * Fo) + (int64_t)c * Fo + fo)
                                   : bw_ld<TG>(dy + e);
        for (int t = t0; t < t1; ++t)
            for (int f = f0; f < f1; ++f)
                bw_st<T>(dz + (((int64_t)b * Tn + t) * F + f) * C + c, (t == at && f == af && best > 0.f) ? g : 0.f);
    }
}

// bf16, channels-last dy, C % 8 == 0: one thread = 8 channels of one pooling window (16-byte accesses)
__global__ void __launch_bounds__(256) maxpool2d_relu_bwd_vec_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ dy,
                                                                     __nv_bfloat16* __restrict__ dz, int B, int Tn, int F, int C,
                                                                     int pt, int pf) {
    pdl_entry();
    const int To = (Tn + pt - 1) / pt, Fo = (F + pf - 1) / pf, C8 = C / 8;
    const int64_t n = (int64_t)B * To * Fo * C8;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
        const int c8 = (int)(e % C8);
        int64_t r = e / C8;
        const int fo = (int)(r % Fo); r /= Fo;
        const int to = (int)(r % To), b = (int)(r / To);
        const int t0 = to * pt, t1 = min(Tn, t0 + pt), f0 = fo * pf, f1 = min(F, f0 + pf);
        float best[8], g[8];
        int arg[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; arg[k] = 0; }
        int w = 0;
        for (int t = t0; t < t1; ++t)
            for (int f = f0; f < f1; ++f, ++w) {
                float v[8];
                unpack8(*reinterpret_cast<const uint4*>(a + (((int64_t)b * Tn + t) * F + f) * C + c8 * 8), v);
#pragma unroll
                for (int k = 0; k < 8; ++k) if (v[k] > best[k]) { best[k] = v[k]; arg[k] = w; }
            }
        unpack8(*reinterpret_cast<const uint4*>(dy + (((int64_t)b * To + to) * Fo + fo) * C + c8 * 8), g);
        w = 0;
        for (int t = t0; t < t1; ++t)
            for (int f = f0; f < f1; ++f, ++w) {
                float o[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = (arg[k] == w && best[k] > 0.f) ? g[k] : 0.f;
                *reinterpret_cast<uint4*>(dz + (((int64_t)b * Tn + t) * F + f) * C + c8 * 8) = pack8(o);
            }
    }
}

unsigned bw_grid(int64_t n) {
    int64_t b = ceil_div64(n, 256);
    int64_t cap = (int64_t)num_sms() * 16;
    return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace
}  // namespace nsp

using namespace nsp;

extern "C" nsp_status nsp_layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* gamma,
                                        float eps, float in_scale, const float* dres, int64_t lddr,
                                        float* dx, int64_t lddx, void* dx_bf16, int64_t lddxb,
                                        float* dgamma, float* dbeta, float* dcol, float dcol_alpha, int M, int D, void* stream) {
    NSP_CHECK_ARG(dy && x && gamma && (dx || dx_bf16), "layernorm_bwd: null pointer");
    NSP_CHECK_ARG(M > 0 && D > 0, "layernorm_bwd: bad shape M=%d D=%d", M, D);
    cudaStream_t st = (cudaStream_t)stream;
    int grid = ceil_div(M, 8);
    const int cap = num_sms() * 4;
    if (grid > cap) grid = cap;
    const bool vec_ok = D % 4 == 0 && D <= 2048 &&
        lddy % 4 == 0 && ldx % 4 == 0 && (!dres || lddr % 4 == 0) && (!dx || lddx % 4 == 0) && (!dx_bf16 || lddxb % 4 == 0) &&
        ((uintptr_t)dy % 16 == 0) && ((uintptr_t)x % 16 == 0) && (!dres || (uintptr_t)dres % 16 == 0) &&
        (!dx || (uintptr_t)dx % 16 == 0) && (!dx_bf16 || (uintptr_t)dx_bf16 % 8 == 0);
    if (!vec_ok) {                 // odd widths / pitches: the scalar kernel
        launch_k(layernorm_bwd_generic_kernel, dim3(grid), dim3(256), 0, st, dy, lddy, x, ldx, gamma, eps, in_scale, dres, lddr, dx, lddx,
                                                           (__nv_bfloat16*)dx_bf16, lddxb, dgamma, dbeta, dcol, dcol_alpha, M, D);
        NSP_LAUNCH_OK();
        return NSP_OK;
    }
    const size_t smem = sizeof(float) * 8 * (size_t)D;
    __nv_bfloat16* dxb = (__nv_bfloat16*)dx_bf16;
#define NSP_LNB(VPT)                                                                                                     \
    do {                                                                                                                 \
        auto kern = layernorm_bwd_kernel<VPT>;                                                                           \
        static size_t attr = 0;                                                                                          \
        if (smem > attr) { NSP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; } \
        launch_k(kern, dim3(grid), dim3(256), smem, st, dy, lddy, x, ldx, gamma, eps, in_scale, dres, lddr, dx, lddx, dxb, lddxb, dgamma, dbeta, dcol, dcol_alpha, M, D); \
    } while (0)
    const int vpt = ceil_div(D / 4, 32);
    if (vpt <= 1) NSP_LNB(1); else if (vpt <= 2) NSP_LNB(2); else if (vpt <= 4) NSP_LNB(4);
    else if (vpt <= 8) NSP_LNB(8); else NSP_LNB(16);
#undef NSP_LNB
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_act_bwd(int is_bf16, int act, const void* dh, const void* z, void* dz, int64_t n, void* stream) {
    NSP_CHECK_ARG(dh && z && dz && n >= 0 && act >= 0 && act <= 4, "act_bwd: bad arguments");
    if (n == 0) return NSP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool al16 = (((uintptr_t)dh | (uintptr_t)z | (uintptr_t)dz) & 15) == 0;
    if (is_bf16 && al16 && n % 8 == 0)
        launch_k(act_bwd_vec_kernel, dim3(bw_grid(n / 8)), dim3(256), 0, st, (const uint4*)dh, (const uint4*)z, (uint4*)dz, n / 8, act);
    else if (is_bf16) launch_k(act_bwd_kernel<__nv_bfloat16>, dim3(bw_grid(n)), dim3(256), 0, st, (const __nv_bfloat16*)dh, (const __nv_bfloat16*)z, (__nv_bfloat16*)dz, n, act);
    else launch_k(act_bwd_kernel<float>, dim3(bw_grid(n)), dim3(256), 0, st, (const float*)dh, (const float*)z, (float*)dz, n, act);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_glu_bwd(int is_bf16, const void* dg, const void* pre, void* dpre, int64_t M, int d, void* stream) {
    NSP_CHECK_ARG(dg && pre && dpre && M > 0 && d > 0, "glu_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const bool al16 = (((uintptr_t)dg | (uintptr_t)pre | (uintptr_t)dpre) & 15) == 0;
    if (is_bf16 && al16 && d % 8 == 0)
        launch_k(glu_bwd_vec_kernel, dim3(bw_grid(M * (d / 8))), dim3(256), 0, st, (const uint4*)dg, (const uint4*)pre, (uint4*)dpre, M, d / 8);
    else if (is_bf16) launch_k(glu_bwd_kernel<__nv_bfloat16>, dim3(bw_grid(M * d)), dim3(256), 0, st, (const __nv_bfloat16*)dg, (const __nv_bfloat16*)pre, (__nv_bfloat16*)dpre, M, d);
    else launch_k(glu_bwd_kernel<float>, dim3(bw_grid(M * d)), dim3(256), 0, st, (const float*)dg, (const float*)pre, (float*)dpre, M, d);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_colsum_acc(int is_bf16, const void* x, int64_t ldx, int M, int N, float alpha, float* y, void* stream) {
    NSP_CHECK_ARG(x && y && M > 0 && N > 0, "colsum_acc: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int vec = is_bf16 ? 8 : 4;
    if (N % vec == 0 && ldx % vec == 0 && ((uintptr_t)x % 16 == 0)) {
        const int cblocks = ceil_div(N, 32 * vec);
        int slices = ceil_div(M, 64);
        const int cap = ceil_div(num_sms() * 4, cblocks);
        if (slices > cap) slices = cap < 1 ? 1 : cap;
        dim3 vgrid((unsigned)cblocks, (unsigned)slices);
        if (is_bf16) launch_k(colsum_acc_vec_kernel<__nv_bfloat16, 8>, dim3(vgrid), dim3(256), 0, st, (const __nv_bfloat16*)x, ldx, y, M, N, alpha);
        else launch_k(colsum_acc_vec_kernel<float, 4>, dim3(vgrid), dim3(256), 0, st, (const float*)x, ldx, y, M, N, alpha);
        NSP_LAUNCH_OK();
        return NSP_OK;
    }
    int slices = ceil_div(M, 256);
    const int cap = ceil_div(num_sms() * 8, ceil_div(N, 32));
    if (slices > cap) slices = cap < 1 ? 1 : cap;
    dim3 grid((unsigned)ceil_div(N, 32), (unsigned)slices);
    if (is_bf16) launch_k(colsum_acc_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)x, ldx, y, M, N, alpha);
    else launch_k(colsum_acc_kernel<float>, dim3(grid), dim3(256), 0, st, (const float*)x, ldx, y, M, N, alpha);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_maxpool_time_bwd(const float* x, const float* dy, float* dx, int B, int T, int D, int factor, void* stream) {
    NSP_CHECK_ARG(x && dy && dx && B > 0 && T > 0 && D > 0 && factor >= 1, "maxpool_time_bwd: bad arguments");
    const int To = (T + factor - 1) / factor;
    launch_k(maxpool_time_bwd_kernel, dim3(bw_grid((int64_t)B * To * D)), dim3(256), 0, (cudaStream_t)stream, x, dy, dx, B, T, D, factor);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_pool_time_bwd(const float* dy, float* dx, int B, int T, int D, int factor, int mode, void* stream) {
    NSP_CHECK_ARG(dy && dx && B > 0 && T > 0 && D > 0 && factor >= 1 && mode >= 1 && mode <= 3, "pool_time_bwd: bad arguments");
    launch_k(pool_time_bwd_kernel, dim3(bw_grid((int64_t)B * T * D)), dim3(256), 0, (cudaStream_t)stream, dy, dx, B, T, D, factor, mode);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_relu_mask(int is_bf16, const void* dx, const void* a, void* dz, int64_t n, void* stream) {
    NSP_CHECK_ARG(dx && a && dz && n >= 0, "relu_mask: bad arguments");
    if (n == 0) return NSP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool al16 = (((uintptr_t)dx | (uintptr_t)a | (uintptr_t)dz) & 15) == 0;
    if (is_bf16 && al16 && n % 8 == 0)
        launch_k(relu_mask_vec_kernel, dim3(bw_grid(n / 8)), dim3(256), 0, st, (const uint4*)dx, (const uint4*)a, (uint4*)dz, n / 8);
    else if (is_bf16) launch_k(relu_mask_kernel<__nv_bfloat16>, dim3(bw_grid(n)), dim3(256), 0, st, (const __nv_bfloat16*)dx, (const __nv_bfloat16*)a, (__nv_bfloat16*)dz, n);
    else launch_k(relu_mask_kernel<float>, dim3(bw_grid(n)), dim3(256), 0, st, (const float*)dx, (const float*)a, (float*)dz, n);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_maxpool2d_relu_bwd(int is_bf16, int dy_bf16, const void* a, const void* dy, void* dz, int B, int T, int F,
                                             int C, int pool_t, int pool_f, int in_chmajor, void* stream) {
    NSP_CHECK_ARG(a && dy && dz && B > 0 && T > 0 && F > 0 && C > 0 && pool_t >= 1 && pool_f >= 1, "maxpool2d_relu_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = (int64_t)B * ceil_div(T, pool_t) * ceil_div(F, pool_f) * C;
    if (is_bf16 && dy_bf16 && !in_chmajor && C % 8 == 0 && (((uintptr_t)a | (uintptr_t)dy | (uintptr_t)dz) & 15) == 0)
        launch_k(maxpool2d_relu_bwd_vec_kernel, dim3(bw_grid(n / 8)), dim3(256), 0, st, (const __nv_bfloat16*)a, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dz, B, T, F, C, pool_t, pool_f);
    else if (is_bf16 && dy_bf16)
        launch_k(maxpool2d_relu_bwd_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(bw_grid(n)), dim3(256), 0, st, (const __nv_bfloat16*)a, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dz, B, T, F, C, pool_t, pool_f, in_chmajor);
    else if (is_bf16)
        launch_k(maxpool2d_relu_bwd_kernel<__nv_bfloat16, float>, dim3(bw_grid(n)), dim3(256), 0, st, (const __nv_bfloat16*)a, (const float*)dy, (__nv_bfloat16*)dz, B, T, F, C, pool_t, pool_f, in_chmajor);
    else if (!dy_bf16)
        launch_k(maxpool2d_relu_bwd_kernel<float, float>, dim3(bw_grid(n)), dim3(256), 0, st, (const float*)a, (const float*)dy, (float*)dz, B, T, F, C, pool_t, pool_f, in_chmajor);
    else { set_error("maxpool2d_relu_bwd: fp32 activations with bf16 gradients are not instantiated"); return NSP_ERR_UNSUPPORTED; }
    NSP_LAUNCH_OK();
    return NSP_OK;
}

// act / GLU backward with the bias gradient fused (bf16 only): mode 0: dz[M,N] = dh * act'(z); mode 1: GLU, dh = dg [M,N/2],
// z = pre [M,N], dz = dpre [M,N].  dbias fp32 [N] is ACCUMULATED.  N % 8 == 0 (mode 1: (N/2) % 8 == 0), dense rows.
extern "C" nsp_status nsp_act_bwd_bias(int mode, int act, const void* dh, const void* z, void* dz, float* dbias, int M, int N,
                                       void* stream) {
    NSP_CHECK_ARG(dh && z && dz && dbias && M > 0 && N > 0, "act_bwd_bias: bad arguments");
    NSP_CHECK_ARG((mode == 0 && N % 8 == 0) || (mode == 1 && N % 16 == 0), "act_bwd_bias: N=%d not vectorisable", N);
    NSP_CHECK_ARG((((uintptr_t)dh | (uintptr_t)z | (uintptr_t)dz) & 15) == 0, "act_bwd_bias: unaligned pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int cblocks = ceil_div(N, 256);
    int slices = ceil_div(M, 64);
    const int cap = ceil_div(num_sms() * 4, cblocks);
    if (slices > cap) slices = cap < 1 ? 1 : cap;
    dim3 grid((unsigned)cblocks, (unsigned)slices);
    if (mode == 0) launch_k(act_bwd_bias_kernel<0>, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)dh, (const __nv_bfloat16*)z, (__nv_bfloat16*)dz, dbias, M, N, act);
    else launch_k(act_bwd_bias_kernel<1>, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)dh, (const __nv_bfloat16*)z, (__nv_bfloat16*)dz, dbias, M, N, act);
    NSP_LAUNCH_OK();
    return NSP_OK;
}
