// Backward of the RNN-T joint network's elementwise / row-wise pieces (the reference: torch autograd through
// decoders/rnn_transducer.py:242 `torch.log_softmax(logits)` and :273 `torch.tanh(self.w_enc(eouts) + self.w_dec(douts))`).
// The GEMM parts (output layer, w_enc, w_dec) run on the tcgen05 dgrad / wgrad kernels; these two are HBM-bound passes:
//
//   nsp_log_softmax_bwd:      dz[r, v] = g * (dlp[r, v] - exp(lp[r, v]) * sum_v dlp[r, v])      in place on dlp
//       lp, dlp fp32 [rows, V]; g = optional DEVICE scalar (the upstream d total / d loss, so that no host sync and no
//       separate scaling pass over the [B, T, U+1, V] tensor is needed).  One CTA per row, row cached in registers.
//   nsp_rnnt_joint_tanh_bwd:  p = dh * (1 - h^2);  de[b, t, :] = sum_u p[b, t, u, :];  dd[b, u, :] = sum_t p[b, t, u, :]
//       h, dh fp32 or bf16 [B, T, U1, J]; de fp32 [B, T, J], dd fp32 [B, U1, J].  Two reduction kernels, each reading h and
//       dh once with J-contiguous (coalesced) accesses; no atomics.
#include "common.cuh"

namespace nsp {
namespace {

constexpr int LSB_VPT = 8;     // values per thread kept in registers: rows up to 256 * 8 = 2048 wide take the fast path

__global__ void __launch_bounds__(256) log_softmax_bwd_kernel(const float* __restrict__ lp, float* __restrict__ dlp, int V,
                                                              const float* __restrict__ gscale) {
    pdl_entry();
    __shared__ float scratch[32];
    const int64_t row = blockIdx.x;
    const float* lr = lp + row * V;
    float* dr = dlp + row * V;
    const float g = gscale ? __ldg(gscale) : 1.f;
    if (V <= 256 * LSB_VPT) {
        float d[LSB_VPT];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < LSB_VPT; ++j) {
            const int i = j * 256 + threadIdx.x;
            d[j] = i < V ? dr[i] : 0.f;
            s += d[j];
        }
        s = block_sum<256>(s, scratch);
#pragma unroll
        for (int j = 0; j < LSB_VPT; ++j) {
            const int i = j * 256 + threadIdx.x;
            if (i < V) dr[i] = g * (d[j] - __expf(__ldg(lr + i)) * s);
        }
    } else {
        float s = 0.f;
        for (int i = threadIdx.x; i < V; i += 256) s += dr[i];
        s = block_sum<256>(s, scratch);
        for (int i = threadIdx.x; i < V; i += 256) dr[i] = g * (dr[i] - __expf(__ldg(lr + i)) * s);
    }
}

template <typename T> __device__ __forceinline__ float ldj(const T* p);
template <> __device__ __forceinline__ float ldj<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldj<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// de[b, t, j] = sum_u dh * (1 - h^2): one thread per (b, t, j), u innermost in the loop (stride J: coalesced over j)
template <typename T>
__global__ void __launch_bounds__(256) joint_bwd_enc_kernel(const T* __restrict__ h, const T* __restrict__ dh,
                                                            float* __restrict__ de, int64_t BT, int U1, int J) {
    pdl_entry();
    const int64_t n = BT * J;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
        const int j = (int)(e % J);
        const int64_t bt = e / J;
        const T* hp = h + bt * U1 * J + j;
        const T* dp = dh + bt * U1 * J + j;
        float acc = 0.f;
        for (int u = 0; u < U1; ++u) {
            const float hv = ldj<T>(hp + (int64_t)u * J);
            acc = fmaf(ldj<T>(dp + (int64_t)u * J), 1.f - hv * hv, acc);
        }
        de[e] = acc;
    }
}

// dd[b, u, j] = sum_t dh * (1 - h^2): one thread per (b, u, j), loop over t (stride U1 * J)
template <typename T>
__global__ void __launch_bounds__(256) joint_bwd_dec_kernel(const T* __restrict__ h, const T* __restrict__ dh,
                                                            float* __restrict__ dd, int B, int Tn, int U1, int J) {
    pdl_entry();
    const int64_t n = (int64_t)B * U1 * J;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
        const int j = (int)(e % J);
        const int64_t bu = e / J;
        const int u = (int)(bu % U1);
        const int64_t b = bu / U1;
        const int64_t base = (b * Tn * U1 + u) * J + j;
        float acc = 0.f;
        for (int t = 0; t < Tn; ++t) {
            const int64_t o = base + (int64_t)t * U1 * J;
            const float hv = ldj<T>(h + o);
            acc = fmaf(ldj<T>(dh + o), 1.f - hv * hv, acc);
        }
        dd[e] = acc;
    }
}

unsigned jb_grid(int64_t n) {
    int64_t b = ceil_div64(n, 256), cap = (int64_t)num_sms() * 16;
    return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace
}  // namespace nsp

using namespace nsp;

extern "C" nsp_status nsp_log_softmax_bwd(const float* lp, float* dlp, int64_t rows, int V, const float* gscale, void* stream) {
    NSP_CHECK_ARG(lp && dlp && rows > 0 && V > 0 && rows < (1ll << 31), "log_softmax_bwd: bad arguments");
    launch_k(log_softmax_bwd_kernel, dim3((unsigned)rows), dim3(256), 0, (cudaStream_t)stream, lp, dlp, V, gscale);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_rnnt_joint_tanh_bwd(int is_bf16, const void* h, const void* dh, float* de, float* dd, int B, int T,
                                              int U1, int J, void* stream) {
    NSP_CHECK_ARG(h && dh && de && dd && B > 0 && T > 0 && U1 > 0 && J > 0, "rnnt_joint_tanh_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t BT = (int64_t)B * T;
    if (is_bf16) {
        launch_k(joint_bwd_enc_kernel<__nv_bfloat16>, dim3(jb_grid(BT * J)), dim3(256), 0, st, (const __nv_bfloat16*)h, (const __nv_bfloat16*)dh, de, BT, U1, J);
        launch_k(joint_bwd_dec_kernel<__nv_bfloat16>, dim3(jb_grid((int64_t)B * U1 * J)), dim3(256), 0, st, (const __nv_bfloat16*)h, (const __nv_bfloat16*)dh, dd, B, T, U1, J);
    } else {
        launch_k(joint_bwd_enc_kernel<float>, dim3(jb_grid(BT * J)), dim3(256), 0, st, (const float*)h, (const float*)dh, de, BT, U1, J);
        launch_k(joint_bwd_dec_kernel<float>, dim3(jb_grid((int64_t)B * U1 * J)), dim3(256), 0, st, (const float*)h, (const float*)dh, dd, B, T, U1, J);
    }
    NSP_LAUNCH_OK();
    return NSP_OK;
}
