// Small streaming kernels: tf32 hi/lo split (parity mode operands), dtype casts.
#include "common.cuh"

namespace nsp {
namespace {

__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                                         float* __restrict__ lo, int64_t n) {
    pdl_entry();
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
    pdl_entry();
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (; i < n; i += stride) y[i] = __float2bfloat16_rn(x[i]);
}

// 8 elements per thread and iteration: two 16-byte loads, one 16-byte store (n % 8 == 0, aligned pointers)
__global__ void __launch_bounds__(256) cast_bf16_vec_kernel(const float4* __restrict__ x, uint4* __restrict__ y, int64_t n8) {
    pdl_entry();
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (int64_t)gridDim.x * 256) {
        const float4 a = x[2 * i], b = x[2 * i + 1];
        __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
        uint4 o;
        o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
        o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
        y[i] = o;
    }
}

__global__ void __launch_bounds__(256) scale_kernel(float* __restrict__ x, float a, int64_t n) {
    pdl_entry();
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (; i < n; i += stride) x[i] *= a;
}

// x[b, t, :] = x[b, t, :] * a + pe[t, :]   (absolute sinusoid positions added to the scaled embedding)
__global__ void __launch_bounds__(256) add_pos_enc_kernel(float* __restrict__ x, const float* __restrict__ pe, float a,
                                                          int64_t n, int64_t TD) {
    pdl_entry();
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (; i < n; i += stride) x[i] = fmaf(x[i], a, __ldg(pe + i % TD));
}

// SpecAugment: zero every element of x [B, T, F] whose frequency bin lies in one of the frequency rectangles or whose frame
// lies in one of the time rectangles (same rectangles for every utterance, frontends/spec_augment.py:112-140).  Write-only.
constexpr int MAXRECT = 32;
struct MaskRects { int nf, nt; int f0[MAXRECT], f1[MAXRECT], t0[MAXRECT], t1[MAXRECT]; };

__global__ void __launch_bounds__(256) mask_rects_kernel(float* __restrict__ x, int64_t n, int T, int F, const MaskRects m) {
    pdl_entry();
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int f = (int)(i % F);
        const int t = (int)((i / F) % T);
        bool hit = false;
        for (int r = 0; r < m.nf; ++r) hit |= (f >= m.f0[r] && f < m.f1[r]);
        for (int r = 0; r < m.nt; ++r) hit |= (t >= m.t0[r] && t < m.t1[r]);
        if (hit) x[i] = 0.f;
    }
}

// column sums of a row-major [M, N] fp32 matrix: block = 32 columns x 8 row slices
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* __restrict__ y, int M, int N) {
    pdl_entry();
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), rs = threadIdx.x >> 5;
    float acc = 0.f;
    if (c < N) for (int r = rs; r < M; r += 8) acc += x[(int64_t)r * N + c];
    part[rs][threadIdx.x & 31] = acc;
    __syncthreads();
    if (rs == 0 && c < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x & 31];
        y[c] = t;
    }
}

// TransformerXL sinusoid table (positional_embedding.py:135-138): row r <-> position -(r+1):
// tab[r, i] = sin(-(r+1) * inv_freq[i]) for i < d/2, cos(-(r+1) * inv_freq[i - d/2]) otherwise.
__global__ void __launch_bounds__(256) xl_pos_table_kernel(const float* __restrict__ inv_freq, float* __restrict__ tab, int rows, int d) {
    pdl_entry();
    const int half = d / 2;
    const int64_t n = (int64_t)rows * d;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
        int r = (int)(e / d), i = (int)(e % d);
        float pos = -(float)(r + 1);
        float ang = pos * inv_freq[i < half ? i : i - half];
        tab[e] = i < half ? sinf(ang) : cosf(ang);
    }
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
    launch_k(split_tf32_kernel, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream, x, hi, lo, n);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream) {
    NSP_CHECK_ARG(x && y && n >= 0, "cast_f32_to_bf16: bad arguments");
    if (n == 0) return NSP_OK;
    if (n % 8 == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0)
        launch_k(cast_bf16_vec_kernel, dim3(ew_grid(n / 8)), dim3(256), 0, (cudaStream_t)stream, (const float4*)x, (uint4*)y, n / 8);
    else
        launch_k(cast_bf16_kernel, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream, x, (__nv_bfloat16*)y, n);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_scale_inplace(float* x, float a, int64_t n, void* stream) {
    NSP_CHECK_ARG(x && n >= 0, "scale_inplace: bad arguments");
    if (n == 0) return NSP_OK;
    launch_k(scale_kernel, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream, x, a, n);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_mask_rects(float* x, int B, int T, int F, const int32_t* freq_rects, int n_freq,
                                     const int32_t* time_rects, int n_time, void* stream) {
    NSP_CHECK_ARG(x && B >= 0 && T >= 0 && F > 0 && n_freq >= 0 && n_time >= 0, "mask_rects: bad arguments");
    NSP_CHECK_ARG((n_freq == 0 || freq_rects) && (n_time == 0 || time_rects), "mask_rects: null rectangle list");
    if (n_freq > MAXRECT || n_time > MAXRECT) { set_error("mask_rects: at most %d masks per axis (got %d, %d)", MAXRECT, n_freq, n_time); return NSP_ERR_UNSUPPORTED; }
    const int64_t n = (int64_t)B * T * F;
    if (n == 0 || (n_freq == 0 && n_time == 0)) return NSP_OK;
    MaskRects m;
    m.nf = n_freq; m.nt = n_time;
    for (int r = 0; r < n_freq; ++r) { m.f0[r] = freq_rects[2 * r]; m.f1[r] = freq_rects[2 * r + 1]; }
    for (int r = 0; r < n_time; ++r) { m.t0[r] = time_rects[2 * r]; m.t1[r] = time_rects[2 * r + 1]; }
    launch_k(mask_rects_kernel, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream, x, n, T, F, m);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_add_pos_enc(float* x, const float* pe, float a, int B, int T, int D, void* stream) {
    NSP_CHECK_ARG(x && pe && B >= 0 && T >= 0 && D > 0, "add_pos_enc: bad arguments");
    const int64_t n = (int64_t)B * T * D;
    if (n == 0) return NSP_OK;
    launch_k(add_pos_enc_kernel, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream, x, pe, a, n, (int64_t)T * D);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_colsum(const float* x, float* y, int M, int N, void* stream) {
    NSP_CHECK_ARG(x && y && M > 0 && N > 0, "colsum: bad arguments");
    launch_k(colsum_kernel, dim3((unsigned)ceil_div(N, 32)), dim3(256), 0, (cudaStream_t)stream, x, y, M, N);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_xl_pos_table(const float* inv_freq, float* table, int rows, int d, void* stream) {
    NSP_CHECK_ARG(inv_freq && table && rows > 0 && d > 0 && d % 2 == 0, "xl_pos_table: bad arguments");
    launch_k(xl_pos_table_kernel, dim3(ew_grid((int64_t)rows * d)), dim3(256), 0, (cudaStream_t)stream, inv_freq, table, rows, d);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

// RNN-T joint pre-activation (decoders/rnn_transducer.py:272-275): out[b,t,u,:] = tanh(enc[b,t,:] + dec[b,u,:])
namespace nsp {
namespace {
template <typename TO>
__global__ void __launch_bounds__(256) joint_tanh_kernel(const float* __restrict__ enc, const float* __restrict__ dec,
                                                         TO* __restrict__ out, int B, int T, int U1, int J) {
    pdl_entry();
    const int64_t n4 = (int64_t)B * T * U1 * (J / 4);
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n4; e += (int64_t)gridDim.x * 256) {
        const int j4 = (int)(e % (J / 4));
        int64_t r = e / (J / 4);
        const int u = (int)(r % U1); r /= U1;
        const int t = (int)(r % T);
        const int b = (int)(r / T);
        const float4 a = __ldg(reinterpret_cast<const float4*>(enc + ((int64_t)b * T + t) * J) + j4);
        const float4 d = __ldg(reinterpret_cast<const float4*>(dec + ((int64_t)b * U1 + u) * J) + j4);
        float v0 = tanhf(a.x + d.x), v1 = tanhf(a.y + d.y), v2 = tanhf(a.z + d.z), v3 = tanhf(a.w + d.w);
        if constexpr (sizeof(TO) == 4) {
            reinterpret_cast<float4*>(out)[e] = make_float4(v0, v1, v2, v3);
        } else {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(v0, v1), p1 = __floats2bfloat162_rn(v2, v3);
            uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
            reinterpret_cast<uint2*>(out)[e] = pk;
        }
    }
}
}  // namespace
}  // namespace nsp

extern "C" nsp_status nsp_rnnt_joint_tanh(const float* enc, const float* dec, void* out, int out_bf16, int B, int T, int U1,
                                          int J, void* stream) {
    NSP_CHECK_ARG(enc && dec && out && B > 0 && T > 0 && U1 > 0 && J > 0 && J % 4 == 0, "rnnt_joint_tanh: bad arguments");
    const int64_t n4 = (int64_t)B * T * U1 * (J / 4);
    if (out_bf16) launch_k(nsp::joint_tanh_kernel<__nv_bfloat16>, dim3(ew_grid(n4)), dim3(256), 0, (cudaStream_t)stream, enc, dec, (__nv_bfloat16*)out, B, T, U1, J);
    else launch_k(nsp::joint_tanh_kernel<float>, dim3(ew_grid(n4)), dim3(256), 0, (cudaStream_t)stream, enc, dec, (float*)out, B, T, U1, J);
    NSP_LAUNCH_OK();
    return NSP_OK;
}
