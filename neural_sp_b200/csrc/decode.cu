// Row-wise softmax / log-softmax, greedy CTC path (argmax + collapse) and its trigger points.
//
// Replaces  CTC.probs / CTC.scores        decoders/ctc.py:197-217   (softmax / log_softmax of output(eouts) / temperature)
//           CTC.greedy                    decoders/ctc.py:219-243   (argmax path, collapse repeats, drop blanks)
//           CTC.trigger_points            decoders/ctc.py:152-195   (first frame of every non-blank run of the argmax path)
//           torch.log_softmax(logits, -1) decoders/rnn_transducer.py:242  (RNN-T lattice input, [B,T,U+1,V] rows)
// HBM-bound: each row is read once into registers (128-bit loads) and written once.
#include "common.cuh"

namespace nsp {
namespace {

// G threads per row (32: warp per row, 256: CTA per row); VPT float4/scalars per thread
template <int G, int VPT, int VEC>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           int64_t rows, int V, int log_mode, float inv_temp) {
    pdl_entry();
    __shared__ float scratch[32];
    constexpr int RPB = 256 / G;
    const int lane = threadIdx.x % G;
    const int64_t row = (int64_t)blockIdx.x * RPB + threadIdx.x / G;
    if (row >= rows) return;
    const float* xr = x + row * V;
    float* yr = y + row * V;
    float v[VPT][VEC];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        const int idx = (j * G + lane) * VEC;
        if (idx < V) {
            if constexpr (VEC == 4) {
                float4 t = ld_stream_f4(xr + idx);
                v[j][0] = t.x * inv_temp; v[j][1] = t.y * inv_temp; v[j][2] = t.z * inv_temp; v[j][3] = t.w * inv_temp;
            } else {
                v[j][0] = __ldg(xr + idx) * inv_temp;
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) m = fmaxf(m, v[j][k]);
        } else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) v[j][k] = -INFINITY;
        }
    }
    if constexpr (G == 32) m = warp_max(m); else m = block_max<256>(m, scratch);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j)
#pragma unroll
        for (int k = 0; k < VEC; ++k) s += __expf(v[j][k] - m);
    if constexpr (G == 32) s = warp_sum(s); else s = block_sum<256>(s, scratch);
    const float lse = m + __logf(s);
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        const int idx = (j * G + lane) * VEC;
        if (idx < V) {
            float o[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) { float lp = v[j][k] - lse; o[k] = log_mode ? lp : __expf(lp); }
            if constexpr (VEC == 4) st_stream_f4(yr + idx, make_float4(o[0], o[1], o[2], o[3]));
            else yr[idx] = o[0];
        }
    }
}

// generic fallback: any V, three passes over L1/L2-resident row
__global__ void __launch_bounds__(256) softmax_rows_generic_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                   int64_t rows, int V, int log_mode, float inv_temp) {
    pdl_entry();
    __shared__ float scratch[32];
    const int64_t row = blockIdx.x;
    const float* xr = x + row * V;
    float* yr = y + row * V;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < V; i += 256) m = fmaxf(m, xr[i] * inv_temp);
    m = block_max<256>(m, scratch);
    float s = 0.f;
    for (int i = threadIdx.x; i < V; i += 256) s += __expf(xr[i] * inv_temp - m);
    s = block_sum<256>(s, scratch);
    const float lse = m + __logf(s);
    for (int i = threadIdx.x; i < V; i += 256) { float lp = xr[i] * inv_temp - lse; yr[i] = log_mode ? lp : __expf(lp); }
}

// first-max argmax per row (warp per row)
__global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ x, int32_t* __restrict__ best, int64_t rows, int V) {
    pdl_entry();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * V;
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int i = lane; i < V; i += 32) { float v = __ldg(xr + i); if (v > bv) { bv = v; bi = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) best[row] = (bi == 0x7fffffff) ? 0 : bi;
}

// per utterance: collapse repeats, drop blanks, record the first frame of every non-blank run
__global__ void ctc_collapse_kernel(const int32_t* __restrict__ best, const int32_t* __restrict__ elens, int B, int T, int blank,
                                    int32_t* __restrict__ hyp, int32_t* __restrict__ hyp_lens, int32_t* __restrict__ trig) {
    pdl_entry();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int Tb = min(max(elens[b], 0), T);
    const int32_t* p = best + (int64_t)b * T;
    int n = 0, prev = -1;
    for (int t = 0; t < Tb; ++t) {
        const int tok = p[t];
        if (tok != blank && (t == 0 || tok != prev)) {
            hyp[(int64_t)b * T + n] = tok;
            trig[(int64_t)b * T + n] = t;
            ++n;
        }
        prev = tok;
    }
    hyp_lens[b] = n;
}

}  // namespace
}  // namespace nsp

using namespace nsp;

extern "C" nsp_status nsp_softmax_rows(const float* x, float* y, int64_t rows, int V, int log_mode, float temperature, void* stream) {
    NSP_CHECK_ARG(x && y && rows > 0 && V > 0 && temperature > 0.f, "softmax_rows: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const float it = 1.f / temperature;
    const bool vec4 = (V % 4 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0);
    const int vec = vec4 ? 4 : 1;
    const int nvec = V / vec;
#define NSP_SM(G, VPT, VEC) launch_k(softmax_rows_kernel<G, VPT, VEC>, dim3((unsigned)ceil_div64(rows, 256 / G)), dim3(256), 0, st, x, y, rows, V, log_mode, it)
    if (nvec <= 32 * 8) {
        const int vpt = ceil_div(nvec, 32);
        if (vec4) { if (vpt <= 1) NSP_SM(32, 1, 4); else if (vpt <= 2) NSP_SM(32, 2, 4); else if (vpt <= 4) NSP_SM(32, 4, 4); else NSP_SM(32, 8, 4); }
        else      { if (vpt <= 1) NSP_SM(32, 1, 1); else if (vpt <= 2) NSP_SM(32, 2, 1); else if (vpt <= 4) NSP_SM(32, 4, 1); else NSP_SM(32, 8, 1); }
    } else if (nvec <= 256 * 12) {
        const int vpt = ceil_div(nvec, 256);
        if (vec4) { if (vpt <= 2) NSP_SM(256, 2, 4); else if (vpt <= 4) NSP_SM(256, 4, 4); else if (vpt <= 8) NSP_SM(256, 8, 4); else NSP_SM(256, 12, 4); }
        else      { if (vpt <= 2) NSP_SM(256, 2, 1); else if (vpt <= 4) NSP_SM(256, 4, 1); else if (vpt <= 8) NSP_SM(256, 8, 1); else NSP_SM(256, 12, 1); }
    } else {
        launch_k(softmax_rows_generic_kernel, dim3((unsigned)rows), dim3(256), 0, st, x, y, rows, V, log_mode, it);
    }
#undef NSP_SM
    NSP_LAUNCH_OK();
    return NSP_OK;
}

extern "C" nsp_status nsp_ctc_greedy(const float* logits, int B, int T, int V, const int32_t* elens, int blank,
                                     int32_t* best, int32_t* hyp, int32_t* hyp_lens, int32_t* trigger, void* stream) {
    NSP_CHECK_ARG(logits && elens && best && hyp && hyp_lens && trigger, "ctc_greedy: null pointer");
    NSP_CHECK_ARG(B > 0 && T > 0 && V > 0, "ctc_greedy: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = (int64_t)B * T;
    launch_k(argmax_rows_kernel, dim3((unsigned)ceil_div64(rows, 8)), dim3(256), 0, st, logits, best, rows, V);
    NSP_LAUNCH_OK();
    launch_k(ctc_collapse_kernel, dim3((unsigned)ceil_div(B, 64)), dim3(64), 0, st, best, elens, B, T, blank, hyp, hyp_lens, trigger);
    NSP_LAUNCH_OK();
    return NSP_OK;
}
