// RNN-Transducer loss, forward + backward (HBM-bound), at the reference's op boundary.
//
// Replaces  warp_rnnt.rnnt_loss(log_probs, labels, frames_lengths, labels_lengths, average_frames=False,
//           reduction='mean', gather=False)   called at  decoders/rnn_transducer.py:248-252
//           (and warprnnt_pytorch.RNNTLoss on CPU, :254-256).  Neither package is vendored in the
//           reference tree; the arithmetic restated here is Graves 2012 eq. 16-20 in the log domain:
//   alpha(t,u) = lse(alpha(t-1,u) + lp[t-1,u,blank], alpha(t,u-1) + lp[t,u-1,y_u]),  alpha(0,0) = 0
//   beta(t,u)  = lse(beta(t+1,u) + lp[t,u,blank],   beta(t,u+1) + lp[t,u,y_{u+1}]),  beta(T-1,U) = lp[T-1,U,blank]
//   nll = -beta(0,0);   d nll / d lp[t,u,blank] = -exp(alpha(t,u) + lp + beta(t+1,u) + nll)   (t = T-1, u = U: beta -> 0)
//                       d nll / d lp[t,u,y_{u+1}] = -exp(alpha(t,u) + lp + beta(t,u+1) + nll);  all other entries 0.
//
// Kernels: K1 gathers the two needed log-probs per lattice cell (the rest of the [B,T,U+1,V] tensor is never
// read); K2 = one CTA per utterance, alpha half / beta half sweep anti-diagonals concurrently (thread per u,
// own previous value in a register, neighbour through double-buffered smem, emission prefetch ring);
// K3 streams the dense gradient (zeros + two values per cell) with 128-bit stores -- the only O(B T U V) traffic.
#include "common.cuh"

namespace nsp {
namespace {

struct RnntParams {
    const float* lp;            // [B,T,U1,V]
    int B, T, U1, V;
    const int32_t* labels;      // [B, U1-1]
    const int32_t* flens; const int32_t* ylens;
    int blank;
    float* nll; float* loss; float* grad;
    float* bl; float* lb;       // [B,T,U1] gathered blank / label log-probs
    float* alpha; float* beta;  // [B,T,U1]
};

__global__ void __launch_bounds__(256) rnnt_gather_kernel(RnntParams p) {
    pdl_entry();
    const int64_t n = (int64_t)p.B * p.T * p.U1;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
        const int u = (int)(e % p.U1);
        const int64_t bt = e / p.U1;
        const int b = (int)(bt / p.T), t = (int)(bt % p.T);
        const int Tb = min(max(p.flens[b], 0), p.T), Ub = min(max(p.ylens[b], 0), p.U1 - 1);
        float vb = 0.f, vl = 0.f;
        if (t < Tb && u <= Ub) {
            const float* row = p.lp + e * (int64_t)p.V;
            vb = __ldg(row + p.blank);
            if (u < Ub) {
                int y = min(max(p.labels[(int64_t)b * (p.U1 - 1) + u], 0), p.V - 1);
                vl = __ldg(row + y);
            }
        }
        p.bl[e] = vb; p.lb[e] = vl;
    }
}

__global__ void __launch_bounds__(1024) rnnt_lattice_kernel(RnntParams p, int HALF) {
    pdl_entry();
    extern __shared__ float sm[];      // alpha: [2][U1], beta: [2][U1]
    __shared__ float s_nll;
    const int b = blockIdx.x, tid = threadIdx.x;
    const bool is_b = tid >= HALF;
    const int j = is_b ? tid - HALF : tid;            // alpha: u = j ; beta: u = Ub - j
    const int U1 = p.U1;
    const int Tb = min(max(p.flens[b], 0), p.T), Ub = min(max(p.ylens[b], 0), U1 - 1);
    float* nb = sm + (is_b ? 2 * U1 : 0);
    const int64_t base = (int64_t)b * p.T * U1;
    const float* bl = p.bl + base;
    const float* lb = p.lb + base;
    float* out = (is_b ? p.beta : p.alpha) + base;
    const bool active = j <= Ub;
    const int u = is_b ? Ub - j : j;
    const int nsteps = Tb + Ub;                       // anti-diagonals 0 .. Tb+Ub-1
    if (Tb <= 0) { if (tid == 0) p.nll[b] = 0.f; return; }

    // emission needed at this thread's k-th cell (k = 0.. Tb-1), its "time-like" index:
    //   alpha: t = k      -> e_own = bl[t-1,u] (t>0), e_nb = lb[t,u-1] (u>0)
    //   beta : t = Tb-1-k -> e_own = bl[t,u] (used with beta(t+1,u)), e_nb = lb[t,u] (used with beta(t,u+1))
    constexpr int PF = 4;
    auto ld_own = [&](int k) -> float {
        if (!active || k < 0 || k >= Tb) return 0.f;
        if (!is_b) return k > 0 ? bl[(int64_t)(k - 1) * U1 + u] : 0.f;
        return bl[(int64_t)(Tb - 1 - k) * U1 + u];
    };
    auto ld_nb = [&](int k) -> float {
        if (!active || k < 0 || k >= Tb) return 0.f;
        if (!is_b) return u > 0 ? lb[(int64_t)k * U1 + (u - 1)] : 0.f;
        return u < Ub ? lb[(int64_t)(Tb - 1 - k) * U1 + u] : 0.f;
    };
    float ro[PF], rn[PF];
#pragma unroll
    for (int q = 0; q < PF; ++q) { ro[q] = ld_own(q); rn[q] = ld_nb(q); }

    float own = NSP_NEG_BIG;
    int k = 0;                                        // this thread's next cell index (starts at step n = j)
    for (int n0 = 0; n0 < nsteps; n0 += PF) {
        float no[PF], nn_[PF];
        // cells processed in this chunk by this thread: k .. k+PF-1 (if the thread has started); prefetch the next ones
        const int kbase = max(n0 - j, 0);
#pragma unroll
        for (int q = 0; q < PF; ++q) { no[q] = ld_own(kbase + PF + q); nn_[q] = ld_nb(kbase + PF + q); }
#pragma unroll
        for (int q = 0; q < PF; ++q) {
            const int n = n0 + q;
            if (n < nsteps) {
                const int kk = n - j;                 // cell index of this thread at step n
                if (active && kk >= 0 && kk < Tb) {
                    // ring slot: cells kbase..kbase+PF-1 are in ro/rn
                    const int slot = kk - kbase;
                    float eo = 0.f, en = 0.f;
#pragma unroll
                    for (int s = 0; s < PF; ++s) if (s == slot) { eo = ro[s]; en = rn[s]; }
                    if (slot >= PF) { eo = ld_own(kk); en = ld_nb(kk); }      // (only when j > n0: first chunk of a late starter)
                    float v;
                    const float left = (j > 0) ? nb[((n - 1) & 1) * U1 + (j - 1)] : NSP_NEG_BIG;
                    if (!is_b) {
                        if (kk == 0 && j == 0) v = 0.f;
                        else v = lse2(kk > 0 ? own + eo : NSP_NEG_BIG, j > 0 ? left + en : NSP_NEG_BIG);
                    } else {
                        if (kk == 0 && j == 0) v = eo;                        // beta(T-1,U) = blank(T-1,U)
                        else v = lse2(kk > 0 ? own + eo : NSP_NEG_BIG, j > 0 ? left + en : NSP_NEG_BIG);
                    }
                    v = fmaxf(v, NSP_NEG_BIG);
                    own = v;
                    nb[(n & 1) * U1 + j] = v;
                    const int t = is_b ? Tb - 1 - kk : kk;
                    out[(int64_t)t * U1 + u] = v;
                    k = kk + 1;
                }
                __syncthreads();
            }
        }
        // rotate the ring only for threads whose window advanced with this chunk
        const int kbase_next = max(n0 + PF - j, 0);
        if (kbase_next == kbase + PF) {
#pragma unroll
            for (int q = 0; q < PF; ++q) { ro[q] = no[q]; rn[q] = nn_[q]; }
        } else {
#pragma unroll
            for (int q = 0; q < PF; ++q) { ro[q] = ld_own(kbase_next + q); rn[q] = ld_nb(kbase_next + q); }
        }
    }
    (void)k;
    if (is_b && j == Ub) {          // this thread owns u = 0; its last cell is t = 0 -> beta(0,0)
        s_nll = -own;
    }
    __syncthreads();
    if (tid == 0) {
        float v = s_nll;
        p.nll[b] = (v < 1.0e29f) ? v : 0.f;          // infeasible (cannot happen for T >= 1) -> 0
    }
}

// dense gradient: one warp per lattice cell row of V entries
template <int VEC>
__global__ void __launch_bounds__(256) rnnt_grad_kernel(RnntParams p) {
    pdl_entry();
    const int lane = threadIdx.x & 31;
    const int64_t cell = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t ncell = (int64_t)p.B * p.T * p.U1;
    if (cell >= ncell) return;
    const int u = (int)(cell % p.U1);
    const int64_t bt = cell / p.U1;
    const int b = (int)(bt / p.T), t = (int)(bt % p.T);
    const int Tb = min(max(p.flens[b], 0), p.T), Ub = min(max(p.ylens[b], 0), p.U1 - 1);
    float gb = 0.f, gl = 0.f;
    int y = -1;
    if (t < Tb && u <= Ub) {
        const float nll = p.nll[b];
        const int64_t base = (int64_t)b * p.T * p.U1;
        const float a = p.alpha[cell];
        const float scale = 1.f / (float)p.B;
        if (t < Tb - 1) gb = -scale * __expf(a + p.bl[cell] + p.beta[base + (int64_t)(t + 1) * p.U1 + u] + nll);
        else if (u == Ub) gb = -scale * __expf(a + p.bl[cell] + nll);
        if (u < Ub) {
            y = min(max(p.labels[(int64_t)b * (p.U1 - 1) + u], 0), p.V - 1);
            gl = -scale * __expf(a + p.lb[cell] + p.beta[cell + 1] + nll);
        }
    }
    if (y == p.blank) { gb += gl; y = -1; }
    float* row = p.grad + cell * (int64_t)p.V;
    if constexpr (VEC == 4) {
        for (int i = lane * 4; i < p.V; i += 128) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.blank >= i && p.blank < i + 4) (&v.x)[p.blank - i] = gb;
            if (y >= i && y < i + 4) (&v.x)[y - i] = gl;
            st_stream_f4(row + i, v);
        }
    } else {
        for (int i = lane; i < p.V; i += 32) row[i] = (i == p.blank) ? gb : ((i == y) ? gl : 0.f);
    }
}

// Training from logits (the joint network's output layer): d loss / d logits in ONE pass over the [B,T,U+1,V] tensor.
// With gb, gl the two non-zero entries of d loss / d log_probs of a cell (as in rnnt_grad_kernel), the log-softmax
// backward collapses to   dz[v] = g * ( gb [v == blank] + gl [v == y] - exp(lp[v]) (gb + gl) ):
// lp is read once and dz written once (fp32, optionally in place over lp, or bf16 = the operand of the output layer's
// dgrad / wgrad GEMMs) -- the dense d loss / d log_probs tensor is never materialised.  g = device scalar (upstream
// gradient of the loss) or null.  Warp per lattice cell.
template <int VEC, typename TO>
__global__ void __launch_bounds__(256) rnnt_grad_logits_kernel(RnntParams p, const float* __restrict__ gscale, TO* dz) {
    pdl_entry();
    const int lane = threadIdx.x & 31;
    const int64_t cell = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t ncell = (int64_t)p.B * p.T * p.U1;
    if (cell >= ncell) return;
    const int u = (int)(cell % p.U1);
    const int64_t bt = cell / p.U1;
    const int b = (int)(bt / p.T), t = (int)(bt % p.T);
    const int Tb = min(max(p.flens[b], 0), p.T), Ub = min(max(p.ylens[b], 0), p.U1 - 1);
    float gb = 0.f, gl = 0.f;
    int y = -1;
    if (t < Tb && u <= Ub) {
        const float nll = p.nll[b];
        const int64_t base = (int64_t)b * p.T * p.U1;
        const float a = p.alpha[cell];
        const float scale = (gscale ? __ldg(gscale) : 1.f) / (float)p.B;
        if (t < Tb - 1) gb = -scale * __expf(a + p.bl[cell] + p.beta[base + (int64_t)(t + 1) * p.U1 + u] + nll);
        else if (u == Ub) gb = -scale * __expf(a + p.bl[cell] + nll);
        if (u < Ub) {
            y = min(max(p.labels[(int64_t)b * (p.U1 - 1) + u], 0), p.V - 1);
            gl = -scale * __expf(a + p.lb[cell] + p.beta[cell + 1] + nll);
        }
    }
    if (y == p.blank) { gb += gl; y = -1; }
    const float gsum = gb + gl;
    const float* lrow = p.lp + cell * (int64_t)p.V;
    TO* row = dz + cell * (int64_t)p.V;
    if (gsum == 0.f) {                                   // padded cell (or fully underflowed): exact zeros, lp not read
        if constexpr (VEC == 4 && sizeof(TO) == 4) {
            for (int i = lane * 4; i < p.V; i += 128) st_stream_f4(reinterpret_cast<float*>(row) + i, make_float4(0.f, 0.f, 0.f, 0.f));
        } else {
            for (int i = lane; i < p.V; i += 32) {
                if constexpr (sizeof(TO) == 4) row[i] = 0.f; else row[i] = __float2bfloat16_rn(0.f);
            }
        }
        return;
    }
    if constexpr (VEC == 4) {
        for (int i = lane * 4; i < p.V; i += 128) {
            const float4 l = *reinterpret_cast<const float4*>(lrow + i);
            float v[4] = {-__expf(l.x) * gsum, -__expf(l.y) * gsum, -__expf(l.z) * gsum, -__expf(l.w) * gsum};
            if (p.blank >= i && p.blank < i + 4) v[p.blank - i] += gb;
            if (y >= i && y < i + 4) v[y - i] += gl;
            if constexpr (sizeof(TO) == 4) {
                st_stream_f4(reinterpret_cast<float*>(row) + i, make_float4(v[0], v[1], v[2], v[3]));
            } else {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
                uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
                *reinterpret_cast<uint2*>(row + i) = pk;
            }
        }
    } else {
        for (int i = lane; i < p.V; i += 32) {
            float v = -__expf(lrow[i]) * gsum + ((i == p.blank) ? gb : 0.f) + ((i == y) ? gl : 0.f);
            if constexpr (sizeof(TO) == 4) row[i] = v; else row[i] = __float2bfloat16_rn(v);
        }
    }
}

__global__ void rnnt_finalize_kernel(RnntParams p) {
    pdl_entry();
    __shared__ float scratch[32];
    float a = 0.f;
    for (int i = threadIdx.x; i < p.B; i += 256) a += p.nll[i];
    a = block_sum<256>(a, scratch);
    if (threadIdx.x == 0) p.loss[0] = a / (float)p.B;       // reduction='mean' (rnn_transducer.py:251)
}

}  // namespace
}  // namespace nsp

using namespace nsp;

extern "C" size_t nsp_rnnt_loss_workspace_bytes(int B, int T, int U1) {
    if (B <= 0 || T <= 0 || U1 <= 0) return 0;
    return 4 * align_up((size_t)B * T * U1 * sizeof(float), 256);
}

extern "C" nsp_status nsp_rnnt_loss_fwd_bwd(const float* log_probs, int B, int T, int U1, int V,
                                            const int32_t* labels, const int32_t* flens, const int32_t* ylens, int blank,
                                            float* nll, float* loss, float* grad,
                                            void* workspace, size_t workspace_bytes, void* stream) {
    NSP_CHECK_ARG(log_probs && flens && ylens && nll && loss && workspace, "rnnt_loss: null pointer");
    NSP_CHECK_ARG(labels || U1 == 1, "rnnt_loss: labels is null");
    NSP_CHECK_ARG(B > 0 && T > 0 && U1 > 0 && V > 1, "rnnt_loss: bad shape B=%d T=%d U+1=%d V=%d", B, T, U1, V);
    NSP_CHECK_ARG(blank >= 0 && blank < V, "rnnt_loss: blank out of range");
    if (U1 > 512) { set_error("rnnt_loss: U+1=%d unsupported (max 512)", U1); return NSP_ERR_UNSUPPORTED; }
    NSP_CHECK_ARG(workspace_bytes >= nsp_rnnt_loss_workspace_bytes(B, T, U1), "rnnt_loss: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    RnntParams p;
    p.lp = log_probs; p.B = B; p.T = T; p.U1 = U1; p.V = V; p.labels = labels; p.flens = flens; p.ylens = ylens;
    p.blank = blank; p.nll = nll; p.loss = loss; p.grad = grad;
    const size_t lat = align_up((size_t)B * T * U1 * sizeof(float), 256);
    char* w = (char*)workspace;
    p.bl = (float*)w; p.lb = (float*)(w + lat); p.alpha = (float*)(w + 2 * lat); p.beta = (float*)(w + 3 * lat);
    const int64_t ncell = (int64_t)B * T * U1;
    {
        int64_t blocks = ceil_div64(ncell, 256), cap = (int64_t)num_sms() * 16;
        launch_k(rnnt_gather_kernel, dim3((unsigned)(blocks < cap ? blocks : cap)), dim3(256), 0, st, p);
        NSP_LAUNCH_OK();
    }
    const int half = (int)align_up((size_t)U1, 32);
    launch_k(rnnt_lattice_kernel, dim3(B), dim3(2 * half), (size_t)4 * U1 * sizeof(float), st, p, half);
    NSP_LAUNCH_OK();
    launch_k(rnnt_finalize_kernel, dim3(1), dim3(256), 0, st, p);
    NSP_LAUNCH_OK();
    if (grad) {
        const unsigned grid = (unsigned)ceil_div64(ncell, 8);
        if (V % 4 == 0 && ((uintptr_t)grad % 16 == 0)) launch_k(rnnt_grad_kernel<4>, dim3(grid), dim3(256), 0, st, p);
        else launch_k(rnnt_grad_kernel<1>, dim3(grid), dim3(256), 0, st, p);
        NSP_LAUNCH_OK();
    }
    return NSP_OK;
}

extern "C" nsp_status nsp_rnnt_grad_logits(const float* log_probs, int B, int T, int U1, int V, const int32_t* labels,
                                           const int32_t* flens, const int32_t* ylens, int blank, const float* nll,
                                           const void* workspace, size_t workspace_bytes, const float* gscale,
                                           void* dz, int dz_bf16, void* stream) {
    NSP_CHECK_ARG(log_probs && flens && ylens && nll && workspace && dz, "rnnt_grad_logits: null pointer");
    NSP_CHECK_ARG(labels || U1 == 1, "rnnt_grad_logits: labels is null");
    NSP_CHECK_ARG(B > 0 && T > 0 && U1 > 0 && V > 1 && blank >= 0 && blank < V, "rnnt_grad_logits: bad shape");
    NSP_CHECK_ARG(workspace_bytes >= nsp_rnnt_loss_workspace_bytes(B, T, U1), "rnnt_grad_logits: workspace too small");
    NSP_CHECK_ARG(!dz_bf16 || dz != (const void*)log_probs, "rnnt_grad_logits: in place only for fp32 output");
    cudaStream_t st = (cudaStream_t)stream;
    RnntParams p;
    p.lp = log_probs; p.B = B; p.T = T; p.U1 = U1; p.V = V; p.labels = labels; p.flens = flens; p.ylens = ylens;
    p.blank = blank; p.nll = const_cast<float*>(nll); p.loss = nullptr; p.grad = nullptr;
    const size_t lat = align_up((size_t)B * T * U1 * sizeof(float), 256);
    char* w = (char*)const_cast<void*>(workspace);
    p.bl = (float*)w; p.lb = (float*)(w + lat); p.alpha = (float*)(w + 2 * lat); p.beta = (float*)(w + 3 * lat);
    const unsigned grid = (unsigned)ceil_div64((int64_t)B * T * U1, 8);
    const bool vec = V % 4 == 0 && ((uintptr_t)log_probs % 16 == 0) && ((uintptr_t)dz % 16 == 0);
    if (dz_bf16) {
        if (vec) launch_k(rnnt_grad_logits_kernel<4, __nv_bfloat16>, dim3(grid), dim3(256), 0, st, p, gscale, (__nv_bfloat16*)dz);
        else launch_k(rnnt_grad_logits_kernel<1, __nv_bfloat16>, dim3(grid), dim3(256), 0, st, p, gscale, (__nv_bfloat16*)dz);
    } else {
        if (vec) launch_k(rnnt_grad_logits_kernel<4, float>, dim3(grid), dim3(256), 0, st, p, gscale, (float*)dz);
        else launch_k(rnnt_grad_logits_kernel<1, float>, dim3(grid), dim3(256), 0, st, p, gscale, (float*)dz);
    }
    NSP_LAUNCH_OK();
    return NSP_OK;
}
