// LayerNorm over the last dimension (HBM-bound): one warp per row, row held in registers,
// two-pass mean / variance in fp32 (eps = 1e-12 in the reference amplifies cancellation otherwise).
//
// Replaces nn.LayerNorm at  encoders/conformer_block.py:53,58,70,76,80 (norm1..norm5),
// encoders/transformer.py:600 (norm_out), encoders/transformer_block.py (norm1/2).
// Writes fp32 and/or bf16 so the next tensor-core GEMM can consume the bf16 copy directly.
#include "common.cuh"

namespace nsp {
namespace {

template <int VPT, int VEC>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int64_t ldx,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, float in_scale, float* __restrict__ y, int64_t ldy,
                                                        __nv_bfloat16* __restrict__ yb, int64_t ldyb, int M, int D) {
    pdl_entry();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    const float* xr = x + row * ldx;
    float v[VPT][VEC];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        int idx = (j * 32 + lane) * VEC;
        if (idx < D) {
            if constexpr (VEC == 4) {
                float4 t = *reinterpret_cast<const float4*>(xr + idx);
                v[j][0] = t.x * in_scale; v[j][1] = t.y * in_scale; v[j][2] = t.z * in_scale; v[j][3] = t.w * in_scale;
            } else {
                v[j][0] = xr[idx] * in_scale;
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) s += v[j][k];
        } else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) v[j][k] = 0.f;
        }
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        int idx = (j * 32 + lane) * VEC;
        if (idx < D) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) { float d = v[j][k] - mean; q += d * d; }
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        int idx = (j * 32 + lane) * VEC;
        if (idx < D) {
            float o[VEC];
            if constexpr (VEC == 4) {
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + idx));
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + idx));
                o[0] = (v[j][0] - mean) * rstd * g4.x + b4.x; o[1] = (v[j][1] - mean) * rstd * g4.y + b4.y;
                o[2] = (v[j][2] - mean) * rstd * g4.z + b4.z; o[3] = (v[j][3] - mean) * rstd * g4.w + b4.w;
            } else {
#pragma unroll
                for (int k = 0; k < VEC; ++k) o[k] = (v[j][k] - mean) * rstd * __ldg(gamma + idx + k) + __ldg(beta + idx + k);
            }
            if (y) {
                if constexpr (VEC == 4) *reinterpret_cast<float4*>(y + row * ldy + idx) = make_float4(o[0], o[1], o[2], o[3]);
                else y[row * ldy + idx] = o[0];
            }
            if (yb) {
                if constexpr (VEC == 4) {
                    __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b = __floats2bfloat162_rn(o[2], o[3]);
                    uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&a); pk.y = *reinterpret_cast<uint32_t*>(&b);
                    *reinterpret_cast<uint2*>(yb + row * ldyb + idx) = pk;
                } else {
                    yb[row * ldyb + idx] = __float2bfloat16_rn(o[0]);
                }
            }
        }
    }
}

// Any D (rows wider than the register-resident kernel takes, e.g. LayerNorm2D over [C, F] = 2560 of the CNN front-end,
// encoders/conv.py:399-421): one CTA per row, two-pass statistics like the fast kernel, the row is re-read from L1 / L2.
__global__ void __launch_bounds__(256) layernorm_generic_kernel(const float* __restrict__ x, int64_t ldx,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float eps, float in_scale, float* __restrict__ y, int64_t ldy,
                                                                __nv_bfloat16* __restrict__ yb, int64_t ldyb, int D) {
    pdl_entry();
    __shared__ float scratch[32];
    const int64_t row = blockIdx.x;
    const float* xr = x + row * ldx;
    float s = 0.f;
    for (int i = threadIdx.x; i < D; i += 256) s += xr[i] * in_scale;
    const float mean = block_sum<256>(s, scratch) / (float)D;
    float q = 0.f;
    for (int i = threadIdx.x; i < D; i += 256) { const float d = xr[i] * in_scale - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(block_sum<256>(q, scratch) / (float)D + eps);
    for (int i = threadIdx.x; i < D; i += 256) {
        const float o = (xr[i] * in_scale - mean) * rstd * __ldg(gamma + i) + __ldg(beta + i);
        if (y) y[row * ldy + i] = o;
        if (yb) yb[row * ldyb + i] = __float2bfloat16_rn(o);
    }
}

}  // namespace
}  // namespace nsp

using namespace nsp;

extern "C" nsp_status nsp_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                                        float in_scale, float* y, int64_t ldy, void* y_bf16, int64_t ldyb, int M, int D,
                                        void* stream) {
    NSP_CHECK_ARG(x && gamma && beta && (y || y_bf16), "layernorm: null pointer");
    NSP_CHECK_ARG(M > 0 && D > 0, "layernorm: bad shape M=%d D=%d", M, D);
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec4 = (D % 4 == 0) && (ldx % 4 == 0) && (!y || ldy % 4 == 0) && (!y_bf16 || ldyb % 4 == 0) &&
                      ((uintptr_t)x % 16 == 0) && (!y || (uintptr_t)y % 16 == 0) && (!y_bf16 || (uintptr_t)y_bf16 % 8 == 0) &&
                      ((uintptr_t)gamma % 16 == 0) && ((uintptr_t)beta % 16 == 0);
    const unsigned grid = (unsigned)ceil_div(M, 8);
    __nv_bfloat16* yb = (__nv_bfloat16*)y_bf16;
#define NSP_LN(VPT, VEC) launch_k(layernorm_kernel<VPT, VEC>, dim3(grid), dim3(256), 0, st, x, ldx, gamma, beta, eps, in_scale, y, ldy, yb, ldyb, M, D)
    if (vec4) {
        const int vpt = ceil_div(D / 4, 32);
        if (vpt <= 1) NSP_LN(1, 4); else if (vpt <= 2) NSP_LN(2, 4); else if (vpt <= 4) NSP_LN(4, 4);
        else if (vpt <= 8) NSP_LN(8, 4); else if (vpt <= 16) NSP_LN(16, 4);
        else launch_k(layernorm_generic_kernel, dim3((unsigned)M), dim3(256), 0, st, x, ldx, gamma, beta, eps, in_scale, y, ldy, yb, ldyb, D);
    } else {
        const int vpt = ceil_div(D, 32);
        if (vpt <= 1) NSP_LN(1, 1); else if (vpt <= 2) NSP_LN(2, 1); else if (vpt <= 4) NSP_LN(4, 1);
        else if (vpt <= 8) NSP_LN(8, 1); else if (vpt <= 16) NSP_LN(16, 1); else if (vpt <= 32) NSP_LN(32, 1);
        else launch_k(layernorm_generic_kernel, dim3((unsigned)M), dim3(256), 0, st, x, ldx, gamma, beta, eps, in_scale, y, ldy, yb, ldyb, D);
    }
#undef NSP_LN
    NSP_LAUNCH_OK();
    return NSP_OK;
}
