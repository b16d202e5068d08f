// CTC forced aligner for sm_100a (integer output, bit-exact target).
//
// Replaces (reference, /root/reference):
//   CTCForcedAligner.__call__ / align   neural_sp/models/seq2seq/decoders/ctc.py:632-753
//   _computes_transition :610-625, _flip_* :540-607, _label_to_path :534
//
// The reference runs three python loops over T' with ~10 small torch ops each.  Every utterance
// is independent, so here one CTA owns one utterance: half the CTA sweeps the forward lattice,
// the other half sweeps the backward lattice in the reference's "flipped" coordinates, then the
// CTA walks the greedy connected path (gamma) with a block-wide first-max argmax per frame.
// All arithmetic is fp32 in the reference's operation order (logsumexp = log(sum(exp(x-max)))+max,
// cum = (emit + alpha_pre) + beta_pre, float-equality against LOG_0 = -1e10) so that the
// reachability test (ctc.py:717) and the argmax (ctc.py:720) see the same values.
// Frames t >= elens[b] never influence frames t < elens[b] (the sweeps are causal in their own
// direction and gamma is reset to a one-hot every frame), so they are skipped.
#include "common.cuh"

namespace nsp {
namespace {

constexpr float kLog0 = -1e10f;   // ctc.py:29

struct AlignParams {
    const float* logits;   // [B,T,V]
    int B, T, V;
    const int32_t* labels; int Lmax;
    const int32_t* elens; const int32_t* ylens;
    int blank;
    int32_t* trig;         // [B, Lmax+1]
    float* emit;           // [B,T,Sm]  lp[t,b,path[s]]
    float* fa;             // [B,T,Sm]  emit + alpha_pre
    float* fb;             // [B,T,Sm]  beta_pre (original coordinates)
    int32_t* best;         // [B,T]
    int Sm;
};

// path over the padded label row: _label_to_path(pad_list(ys, 0), blank) (ctc.py:534,646)
__device__ __forceinline__ int apath(const int32_t* lab, int L, int s, int blank) {
    if (!(s & 1)) return blank;
    int q = s >> 1;
    return q < L ? lab[q] : 0;
}

__device__ __forceinline__ float lse3_exact(float m0, float m1, float m2) {
    float mx = fmaxf(fmaxf(m0, m1), m2);
    float s = expf(m0 - mx) + expf(m1 - mx);
    s = s + expf(m2 - mx);
    return logf(s) + mx;
}

// one CTA (128 threads) per (b,t) row: fp32 log-softmax statistics + path gather
__global__ void __launch_bounds__(128) align_rows_kernel(AlignParams p) {
    pdl_entry();
    __shared__ float scratch[32];
    const int64_t row = blockIdx.x;
    const int b = (int)(row / p.T), t = (int)(row % p.T);
    const int Tb = min(max(p.elens[b], 0), p.T);
    if (t >= Tb) return;
    const float* x = p.logits + row * (int64_t)p.V;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < p.V; i += 128) m = fmaxf(m, __ldg(x + i));
    m = block_max<128>(m, scratch);
    float s = 0.f;
    for (int i = threadIdx.x; i < p.V; i += 128) s += expf(__ldg(x + i) - m);
    s = block_sum<128>(s, scratch);
    const float ls = logf(s);
    const int L = min(max(p.ylens[b], 0), p.Lmax);
    const int32_t* lab = p.labels + (int64_t)b * p.Lmax;
    float* em = p.emit + row * (int64_t)p.Sm;
    for (int k = threadIdx.x; k < p.Sm; k += 128) {
        int v = min(max(apath(lab, L, k, p.blank), 0), p.V - 1);
        em[k] = (__ldg(x + v) - m) - ls;
    }
}

template <int SPT>
__global__ void __launch_bounds__(1024) align_lattice_kernel(AlignParams p, int HALF) {
    pdl_entry();
    extern __shared__ float sm[];
    const int Sm = p.Sm;
    float* abuf = sm;               // [2][Sm]
    float* bbuf = sm + 2 * Sm;      // [2][Sm]
    __shared__ float r_val[32];
    __shared__ int r_idx[32];
    __shared__ int s_o;

    const int b = blockIdx.x, tid = threadIdx.x;
    const bool is_b = tid >= HALF;
    const int htid = is_b ? tid - HALF : tid;
    const int L = min(max(p.ylens[b], 0), p.Lmax);
    const int Sb = 2 * L + 1;
    const int Tb = min(max(p.elens[b], 0), p.T);
    const int32_t* lab = p.labels + (int64_t)b * p.Lmax;
    const int64_t base = (int64_t)b * p.T * Sm;
    const float* em = p.emit + base;
    float* fa = p.fa + base;
    float* fb = p.fb + base;
    constexpr int PF = 4;

    // static per-state info.  alpha half: original coordinate s.  beta half: flipped coordinate k,
    // sigma(k) = (Sm-1-k+Sb) % Sm is the original state it maps to (ctc.py:540-561,586-607).
    int st[SPT], src[SPT];
    bool valid[SPT], same[SPT], outside[SPT];
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
        int s = htid + j * HALF;
        st[j] = s; valid[j] = s < Sm; src[j] = 0; same[j] = false; outside[j] = true;
        if (valid[j]) {
            outside[j] = s >= Sb;
            if (!is_b) {
                src[j] = s;
                same[j] = s >= 2 && apath(lab, L, s - 2, p.blank) == apath(lab, L, s, p.blank);
            } else {
                src[j] = (Sm - 1 - s + Sb) % Sm;
                if (s >= 2) {
                    int s2 = (Sm - 1 - (s - 2) + Sb) % Sm;
                    same[j] = apath(lab, L, s2, p.blank) == apath(lab, L, src[j], p.blank);
                }
            }
        }
    }
    float* buf = is_b ? bbuf : abuf;
    for (int s = htid; s < Sm; s += HALF) { buf[s] = kLog0; buf[Sm + s] = (s == 0) ? 0.f : kLog0; }
    __syncthreads();

    // ---- alpha (forward) and beta (backward, flipped) sweeps, concurrently ----
    {
        auto tstep = [&](int i) { return is_b ? (Tb - 1 - i) : i; };
        float ering[SPT][PF];
#pragma unroll
        for (int j = 0; j < SPT; ++j)
#pragma unroll
            for (int q = 0; q < PF; ++q)
                ering[j][q] = (valid[j] && q < Tb) ? em[(int64_t)tstep(q) * Sm + src[j]] : 0.f;
        int cur = 0;
        for (int i0 = 0; i0 < Tb; i0 += PF) {
            float enext[SPT][PF];
#pragma unroll
            for (int j = 0; j < SPT; ++j)
#pragma unroll
                for (int q = 0; q < PF; ++q) {
                    int i = i0 + PF + q;
                    enext[j][q] = (valid[j] && i < Tb) ? em[(int64_t)tstep(i) * Sm + src[j]] : 0.f;
                }
#pragma unroll
            for (int q = 0; q < PF; ++q) {
                int i = i0 + q;
                if (i < Tb) {
                    const float* prev = buf + (cur ^ 1) * Sm;
                    float* now = buf + cur * Sm;
                    const int t = tstep(i);
#pragma unroll
                    for (int j = 0; j < SPT; ++j) {
                        if (valid[j]) {
                            int s = st[j];
                            float m0 = prev[s];
                            float m1 = s >= 1 ? prev[s - 1] : kLog0;
                            float m2 = (s >= 2 && !same[j]) ? prev[s - 2] : kLog0;
                            float nw = lse3_exact(m0, m1, m2);
                            if (outside[j]) nw = kLog0;
                            float stt = nw + ering[j][q];
                            if (!is_b) fa[(int64_t)t * Sm + s] = stt;        // emit + alpha_pre (== new state)
                            else       fb[(int64_t)t * Sm + src[j]] = nw;    // beta_pre, original coordinates
                            now[s] = stt;
                        }
                    }
                    __syncthreads();
                    cur ^= 1;
                }
            }
#pragma unroll
            for (int j = 0; j < SPT; ++j)
#pragma unroll
                for (int q = 0; q < PF; ++q) ering[j][q] = enext[j][q];
        }
    }
    __syncthreads();

    // ---- gamma: greedy connected best path (ctc.py:712-729); first HALF*SPT threads own states ----
    float* g = abuf;   // reuse: [2][Sm]
    for (int s = tid; s < Sm; s += blockDim.x) { g[s] = kLog0; g[Sm + s] = (s == 0) ? 0.f : kLog0; }
    __syncthreads();
    // re-derive per-state info in ORIGINAL coordinates for every owning thread
    bool own = tid < HALF;   // alpha-half threads already hold original-coordinate info
    int cur = 0;
    const int lane = tid & 31, wid = tid >> 5, nwarp = blockDim.x >> 5;
    for (int t = 0; t < Tb; ++t) {
        const float* prev = g + (cur ^ 1) * Sm;
        float best_v = -INFINITY; int best_i = 0x7fffffff;
        if (own) {
#pragma unroll
            for (int j = 0; j < SPT; ++j) {
                if (valid[j]) {
                    int s = st[j];
                    float m0 = prev[s];
                    float m1 = s >= 1 ? prev[s - 1] : kLog0;
                    float m2 = (s >= 2 && !same[j]) ? prev[s - 2] : kLog0;
                    float nw = lse3_exact(m0, m1, m2);
                    if (outside[j]) nw = kLog0;
                    float gv = nw + em[(int64_t)t * Sm + s];
                    float v = fa[(int64_t)t * Sm + s] + fb[(int64_t)t * Sm + s];
                    if (gv == kLog0) v = kLog0;
                    if (v > best_v || (v == best_v && s < best_i)) { best_v = v; best_i = s; }
                }
            }
        }
        // block-wide first-max argmax
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
            int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
        }
        if (lane == 0) { r_val[wid] = best_v; r_idx[wid] = best_i; }
        __syncthreads();
        if (wid == 0) {
            float v = lane < nwarp ? r_val[lane] : -INFINITY;
            int ix = lane < nwarp ? r_idx[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float ov = __shfl_xor_sync(0xffffffffu, v, o);
                int oi = __shfl_xor_sync(0xffffffffu, ix, o);
                if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
            }
            if (lane == 0) {
                s_o = ix;
                p.best[(int64_t)b * p.T + t] = apath(lab, L, ix, p.blank);
            }
        }
        __syncthreads();
        const int o = s_o;
        float* now = g + cur * Sm;
        for (int s = tid; s < Sm; s += blockDim.x) now[s] = (s == o) ? 0.f : kLog0;
        __syncthreads();
        cur ^= 1;
    }

    // ---- trigger points (ctc.py:732-750): first frame of every non-blank run; <eos> slot = elens-1 ----
    if (tid == 0) {
        int32_t* tr = p.trig + (int64_t)b * (p.Lmax + 1);
        const int32_t* bs = p.best + (int64_t)b * p.T;
        tr[L] = Tb - 1;
        int n = 0, prevtok = -1;
        for (int t = 0; t < Tb; ++t) {
            int tok = bs[t];
            if (tok != p.blank && (t == 0 || tok != prevtok) && n <= p.Lmax) tr[n++] = t;
            prevtok = tok;
        }
    }
}

}  // namespace
}  // namespace nsp

using namespace nsp;

extern "C" size_t nsp_ctc_align_workspace_bytes(int B, int T, int Lmax) {
    if (B <= 0 || T <= 0 || Lmax < 0) return 0;
    size_t Sm = 2 * (size_t)Lmax + 1, bt = (size_t)B * T;
    return align_up(3 * bt * Sm * sizeof(float), 256) + align_up(bt * sizeof(int32_t), 256);
}

extern "C" nsp_status nsp_ctc_forced_align(const float* logits, int B, int T, int V,
                                           const int32_t* labels, int Lmax,
                                           const int32_t* elens, const int32_t* ylens, int blank,
                                           int32_t* trigger_points,
                                           void* workspace, size_t workspace_bytes, void* stream) {
    NSP_CHECK_ARG(logits && labels && elens && ylens && trigger_points && workspace, "ctc_align: null pointer");
    NSP_CHECK_ARG(B > 0 && T > 0 && V > 1 && Lmax > 0, "ctc_align: bad shape B=%d T=%d V=%d Lmax=%d", B, T, V, Lmax);
    NSP_CHECK_ARG(blank >= 0 && blank < V, "ctc_align: blank out of range");
    if (2 * Lmax + 1 > 2048) { set_error("ctc_align: Lmax=%d unsupported (2L+1 <= 2048)", Lmax); return NSP_ERR_UNSUPPORTED; }
    NSP_CHECK_ARG(workspace_bytes >= nsp_ctc_align_workspace_bytes(B, T, Lmax), "ctc_align: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    AlignParams p;
    p.logits = logits; p.B = B; p.T = T; p.V = V; p.labels = labels; p.Lmax = Lmax;
    p.elens = elens; p.ylens = ylens; p.blank = blank; p.trig = trigger_points;
    p.Sm = 2 * Lmax + 1;
    const size_t bt = (size_t)B * T, lat = bt * p.Sm * sizeof(float);
    char* w = (char*)workspace;
    p.emit = (float*)w; p.fa = (float*)(w + lat); p.fb = (float*)(w + 2 * lat);
    p.best = (int32_t*)(w + align_up(3 * lat, 256));
    NSP_CUDA_OK(cudaMemsetAsync(trigger_points, 0, (size_t)B * (Lmax + 1) * sizeof(int32_t), st));
    launch_k(align_rows_kernel, dim3((unsigned)bt), dim3(128), 0, st, p);
    NSP_LAUNCH_OK();
    int spt = 1, half = (int)align_up((size_t)p.Sm, 32);
    if (half > 512) { spt = 2; half = (int)align_up((size_t)ceil_div(p.Sm, 2), 32); }
    if (half > 512) { spt = 4; half = 512; }
    size_t smem = (size_t)4 * p.Sm * sizeof(float);
    if (spt == 1) launch_k(align_lattice_kernel<1>, dim3(B), dim3(2 * half), smem, st, p, half);
    else if (spt == 2) launch_k(align_lattice_kernel<2>, dim3(B), dim3(2 * half), smem, st, p, half);
    else launch_k(align_lattice_kernel<4>, dim3(B), dim3(2 * half), smem, st, p, half);
    NSP_LAUNCH_OK();
    return NSP_OK;
}
