// CTC loss forward+backward for sm_100a: one streaming pass over the logits, a per-utterance
// alpha/beta lattice sweep, and a sparse fix-up of the gradient.
//
// Replaces (reference, /root/reference):
//   CTC.loss_fn                neural_sp/models/seq2seq/decoders/ctc.py:139-150
//   kldiv_lsm_ctc              neural_sp/models/criterion.py:110-127  (mixed at ctc.py:128-129)
//   and the autograd backward of both (ATen ctc_loss_backward + log_softmax backward).
//
// Roofline: HBM.  Algorithmic traffic = read logits once + write d(loss)/d(logits) once
// = 8 B per logit element.  Kernel plan:
//   K1 ctc_rows_kernel     grid over (b,t) rows: row -> registers (128-bit streaming loads),
//                          max / sum-exp / entropy, writes grad = softmax-part (+ label-smoothing
//                          KL part) with streaming stores, gathers the 2L+1 path emissions.
//                          Rows t >= elens[b] are zero-filled without being read.
//   K2 ctc_lattice_kernel  one CTA per utterance; half the CTA sweeps alpha forward while the
//                          other half sweeps beta backward (one state per thread, smem ping-pong,
//                          register prefetch ring for the emissions), then all warps apply the
//                          sparse "-= occupancy" fix-up to the <= L+1 touched columns of each row.
//   K3 ctc_finalize_kernel deterministic reduction of nll / KL to the scalar loss.
#include "tc_common.cuh"
#include <stdlib.h>

namespace nsp {
namespace {

struct CtcParams {
    const float* logits;
    int64_t sb, st;
    int B, T, V;
    const int32_t* labels;
    int Lmax;
    const int32_t* elens;
    const int32_t* ylens;
    int blank;
    float lsm;
    float* nll;
    float* loss;
    float* grad;
    // workspace
    float* emit;   // [B,T,Sp]
    float* alpha;  // [B,T,Sp]
    float* beta;   // [B,T,Sp]
    float* lse;    // [B,T]
    float* hrow;   // [B,T]  sum_v p*lp
    float* klrow;  // [B,T]
    float* nll_raw;    // [B] un-zeroed nll (>= 1e29 when infeasible)
    int16_t* nxt;      // [B,Sp] next state carrying the same label (-1: none)
    int16_t* head;     // [B,Sp] 1 if first occurrence of its label in the path
    int Sp;            // lattice row pitch (2*Lmax+1 rounded up to a multiple of 16)
};

template <int G>
__device__ __forceinline__ float group_max(float v, float* scratch) {
    if constexpr (G == 32) return warp_max(v);
    else return block_max<G>(v, scratch);
}
template <int G>
__device__ __forceinline__ float group_sum(float v, float* scratch) {
    if constexpr (G == 32) return warp_sum(v);
    else return block_sum<G>(v, scratch);
}

__device__ __forceinline__ int path_label(const int32_t* lab, int s, int blank, int V) {
    int l = (s & 1) ? lab[s >> 1] : blank;
    return min(max(l, 0), V - 1);   // memory safety for out-of-range ids (reference would raise)
}

__device__ __forceinline__ float n_frames_of(const int32_t* elens, int B, int T) {
    // sum_b elens[b] (criterion.py:126 denominator); B is small, L2-resident.
    int acc = 0;
    for (int i = 0; i < B; ++i) acc += min(max(elens[i], 0), T);
    return (float)acc;
}

// ---------------------------------------------------------------------------------------------
// K1, register-resident rows.  G threads cooperate on one row; 256 threads per CTA.
// ---------------------------------------------------------------------------------------------
template <int G, int VPT, int VEC>
__global__ void __launch_bounds__(256, (VPT * VEC <= 32) ? 4 : 3) ctc_rows_kernel(CtcParams p) {
    pdl_entry();
    constexpr int NT = 256;
    constexpr int RPB = NT / G;
    __shared__ float scratch[32];
    const int lane = threadIdx.x % G;
    const int64_t row = (int64_t)blockIdx.x * RPB + threadIdx.x / G;
    if (row >= (int64_t)p.B * p.T) return;   // uniform per group (G==256: per CTA)
    const int b = (int)(row / p.T), t = (int)(row % p.T);
    const int Tb = min(max(p.elens[b], 0), p.T);
    float* grow = p.grad + row * (int64_t)p.V;
    const int V = p.V;

    if (t >= Tb) {   // padded frame: zero gradient, nothing read
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            int idx = (j * G + lane) * VEC;
            if (idx < V) {
                if constexpr (VEC == 4) st_stream_f4(grow + idx, make_float4(0.f, 0.f, 0.f, 0.f));
                else grow[idx] = 0.f;
            }
        }
        if (lane == 0) { p.klrow[row] = 0.f; p.lse[row] = 0.f; p.hrow[row] = 0.f; }
        return;
    }

    const float* xrow = p.logits + (int64_t)b * p.sb + (int64_t)t * p.st;
    float x[VPT][VEC];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        int idx = (j * G + lane) * VEC;
        if (idx < V) {
            if constexpr (VEC == 4) {
                float4 v = ld_stream_f4(xrow + idx);
                x[j][0] = v.x; x[j][1] = v.y; x[j][2] = v.z; x[j][3] = v.w;
            } else {
                x[j][0] = __ldg(xrow + idx);
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) m = fmaxf(m, x[j][k]);
        } else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) x[j][k] = -INFINITY;
        }
    }
    m = group_max<G>(m, scratch);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j)
#pragma unroll
        for (int k = 0; k < VEC; ++k) sum += __expf(x[j][k] - m);   // exp(-inf) = 0 for the tail
    sum = group_sum<G>(sum, scratch);
    const float lse = m + __logf(sum);

    const float c_ctc = (1.f - p.lsm) / (float)p.B;
    float c_kl = 0.f, H = 0.f;
    if (p.lsm > 0.f) {
        float h = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j)
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                float lp = x[j][k] - lse;
                float pv = __expf(lp);
                if (pv > 0.f) h += pv * lp;
            }
        H = group_sum<G>(h, scratch);
        c_kl = p.lsm / n_frames_of(p.elens, p.B, p.T);
    }

    // the exponentials are recomputed (MUFU has headroom at HBM speed) instead of kept: half the registers,
    // twice the resident CTAs, more bytes in flight per SM
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        int idx = (j * G + lane) * VEC;
        if (idx < V) {
            float g[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                float lp = x[j][k] - lse;
                float pv = __expf(lp);
                g[k] = (pv > 0.f) ? pv * (c_ctc + c_kl * (lp - H)) : 0.f;   // p == 0: lp may be -inf
            }
            if constexpr (VEC == 4) st_stream_f4(grow + idx, make_float4(g[0], g[1], g[2], g[3]));
            else grow[idx] = g[0];
        }
    }
    if (lane == 0) {
        p.lse[row] = lse;
        p.hrow[row] = H;
        // sum_v p (lp - log(1/(V-1))) = H + log(V-1)
        p.klrow[row] = (p.lsm > 0.f) ? (H + __logf((float)(V - 1))) : 0.f;
    }
    // path emissions (L2 hits: this row was just streamed by this group)
    const int S = 2 * min(max(p.ylens[b], 0), p.Lmax) + 1;
    const int32_t* lab = p.labels + (int64_t)b * p.Lmax;
    float* em = p.emit + row * (int64_t)p.Sp;
    for (int s = lane; s < S; s += G) em[s] = __ldg(xrow + path_label(lab, s, p.blank, V)) - lse;
}

// K1, generic: row staged in shared memory (any V up to ~50k, any alignment).
template <int NT>
__global__ void __launch_bounds__(NT) ctc_rows_smem_kernel(CtcParams p) {
    pdl_entry();
    extern __shared__ float srow[];
    __shared__ float scratch[32];
    const int64_t row = blockIdx.x;
    const int b = (int)(row / p.T), t = (int)(row % p.T);
    const int Tb = min(max(p.elens[b], 0), p.T);
    float* grow = p.grad + row * (int64_t)p.V;
    const int V = p.V;
    if (t >= Tb) {
        for (int i = threadIdx.x; i < V; i += NT) grow[i] = 0.f;
        if (threadIdx.x == 0) { p.klrow[row] = 0.f; p.lse[row] = 0.f; p.hrow[row] = 0.f; }
        return;
    }
    const float* xrow = p.logits + (int64_t)b * p.sb + (int64_t)t * p.st;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < V; i += NT) { float v = __ldg(xrow + i); srow[i] = v; m = fmaxf(m, v); }
    m = block_max<NT>(m, scratch);
    float sum = 0.f;
    for (int i = threadIdx.x; i < V; i += NT) sum += __expf(srow[i] - m);
    sum = block_sum<NT>(sum, scratch);
    const float lse = m + __logf(sum);
    const float c_ctc = (1.f - p.lsm) / (float)p.B;
    float c_kl = 0.f, H = 0.f;
    if (p.lsm > 0.f) {
        float h = 0.f;
        for (int i = threadIdx.x; i < V; i += NT) {
            float lp = srow[i] - lse; float pv = __expf(lp);
            if (pv > 0.f) h += pv * lp;
        }
        H = block_sum<NT>(h, scratch);
        c_kl = p.lsm / n_frames_of(p.elens, p.B, p.T);
    }
    for (int i = threadIdx.x; i < V; i += NT) {
        float lp = srow[i] - lse; float pv = __expf(lp);
        grow[i] = (pv > 0.f) ? pv * (c_ctc + c_kl * (lp - H)) : 0.f;
    }
    if (threadIdx.x == 0) {
        p.lse[row] = lse; p.hrow[row] = H;
        p.klrow[row] = (p.lsm > 0.f) ? (H + __logf((float)(V - 1))) : 0.f;
    }
    const int S = 2 * min(max(p.ylens[b], 0), p.Lmax) + 1;
    const int32_t* lab = p.labels + (int64_t)b * p.Lmax;
    float* em = p.emit + row * (int64_t)p.Sp;
    for (int s = threadIdx.x; s < S; s += NT) em[s] = srow[path_label(lab, s, p.blank, V)] - lse;
}

// ---------------------------------------------------------------------------------------------
// K2: lattice.  CTA = 2*HALF threads; threads [0,HALF) run alpha forward, [HALF,2*HALF) run beta
// backward, SPT states per thread (state = htid + k*HALF).
// ---------------------------------------------------------------------------------------------
template <int SPT>
__global__ void __launch_bounds__(1024) ctc_lattice_kernel(CtcParams p, int HALF) {
    pdl_entry();
    extern __shared__ float sm[];
    const int Sp = p.Sp;
    float* abuf = sm;                 // [2][Sp] alpha ping-pong
    float* bbuf = sm + 2 * Sp;        // [2][Sp] beta ping-pong
    int16_t* nxt = reinterpret_cast<int16_t*>(sm + 4 * Sp);   // [Sp] next state with the same label (-1: none)
    int16_t* head = nxt + Sp;                                  // [Sp] 1 if first occurrence of its label
    __shared__ float s_nll;

    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const bool is_beta = tid >= HALF;
    const int htid = is_beta ? tid - HALF : tid;
    const int L = min(max(p.ylens[b], 0), p.Lmax);
    const int S = 2 * L + 1;
    const int Tb = min(max(p.elens[b], 0), p.T);
    const int32_t* lab = p.labels + (int64_t)b * p.Lmax;
    const int64_t base = (int64_t)b * p.T * Sp;
    const float* em = p.emit + base;
    float* gout = (is_beta ? p.beta : p.alpha) + base;

    // same-label chains over the odd (label) states: O(L) per state, done once
    for (int s = tid; s < S; s += blockDim.x) {
        int nx = -1, hd = 1;
        if (s & 1) {
            int me = lab[s >> 1];
            for (int q = (s >> 1) + 1; q < L; ++q) if (lab[q] == me) { nx = 2 * q + 1; break; }
            for (int q = 0; q < (s >> 1); ++q) if (lab[q] == me) { hd = 0; break; }
        }
        nxt[s] = (int16_t)nx; head[s] = (int16_t)hd;
    }

    // per-state static info
    bool valid[SPT], skip[SPT];
    int st_[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        int s = htid + k * HALF;
        st_[k] = s;
        valid[k] = s < S;
        skip[k] = false;
        if (valid[k] && (s & 1)) {
            if (!is_beta) skip[k] = (s >= 2) && (lab[s >> 1] != lab[(s >> 1) - 1]);
            else          skip[k] = (s + 2 < S) && (lab[s >> 1] != lab[(s >> 1) + 1]);
        }
    }
    float* buf = is_beta ? bbuf : abuf;
    // init ping-pong buffer 1 ("previous") to NEG so step 0 can use the common recurrence
    for (int s = htid; s < Sp; s += HALF) { buf[s] = NSP_NEG_BIG; buf[Sp + s] = NSP_NEG_BIG; }
    __syncthreads();

    if (Tb > 0) {
        constexpr int PF = 4;   // emission prefetch ring (steps)
        float ering[SPT][PF];
        auto tstep = [&](int i) { return is_beta ? (Tb - 1 - i) : i; };   // i-th step of this sweep
#pragma unroll
        for (int k = 0; k < SPT; ++k)
#pragma unroll
            for (int j = 0; j < PF; ++j)
                ering[k][j] = (valid[k] && j < Tb) ? em[(int64_t)tstep(j) * Sp + st_[k]] : 0.f;

        int cur = 0;
        for (int i0 = 0; i0 < Tb; i0 += PF) {
            float enext[SPT][PF];
#pragma unroll
            for (int k = 0; k < SPT; ++k)
#pragma unroll
                for (int j = 0; j < PF; ++j) {
                    int i = i0 + PF + j;
                    enext[k][j] = (valid[k] && i < Tb) ? em[(int64_t)tstep(i) * Sp + st_[k]] : 0.f;
                }
#pragma unroll
            for (int j = 0; j < PF; ++j) {
                int i = i0 + j;
                if (i < Tb) {   // uniform across the CTA
                    const float* prev = buf + (cur ^ 1) * Sp;
                    float* now = buf + cur * Sp;
                    int t = tstep(i);
#pragma unroll
                    for (int k = 0; k < SPT; ++k) {
                        if (valid[k]) {
                            int s = st_[k];
                            float v;
                            if (i == 0) {
                                bool start = is_beta ? (s >= S - 2) : (s <= 1);
                                v = start ? ering[k][j] : NSP_NEG_BIG;
                            } else {
                                float a0 = prev[s], a1, a2;
                                if (!is_beta) {
                                    a1 = (s >= 1) ? prev[s - 1] : NSP_NEG_BIG;
                                    a2 = skip[k] ? prev[s - 2] : NSP_NEG_BIG;
                                } else {
                                    a1 = (s + 1 < S) ? prev[s + 1] : NSP_NEG_BIG;
                                    a2 = skip[k] ? prev[s + 2] : NSP_NEG_BIG;
                                }
                                v = lse3(a0, a1, a2) + ering[k][j];
                                v = fmaxf(v, NSP_NEG_BIG);
                            }
                            now[s] = v;
                            gout[(int64_t)t * Sp + s] = v;
                        }
                    }
                    __syncthreads();
                    cur ^= 1;
                }
            }
#pragma unroll
            for (int k = 0; k < SPT; ++k)
#pragma unroll
                for (int j = 0; j < PF; ++j) ering[k][j] = enext[k][j];
        }
        if (tid == 0) {
            const float* last = abuf + (cur ^ 1) * Sp;
            float l1 = last[S - 1];
            float l2 = (S > 1) ? last[S - 2] : NSP_NEG_BIG;
            s_nll = -lse2(l1, l2);
        }
    } else if (tid == 0) {
        s_nll = (L == 0) ? 0.f : 1.0e30f;
    }
    __syncthreads();   // also makes the alpha/beta global stores of this CTA visible to itself
    const float nll = s_nll;
    const bool feasible = nll < 1.0e29f;
    if (tid == 0) p.nll[b] = feasible ? nll : 0.f;

    const float c_ctc = (1.f - p.lsm) / (float)p.B;
    float* gb = p.grad + (int64_t)b * p.T * p.V;
    const int lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;

    if (!feasible) {
        // zero_infinity: CTC part of the gradient vanishes; only the label-smoothing KL part stays.
        float c_kl = (p.lsm > 0.f) ? p.lsm / n_frames_of(p.elens, p.B, p.T) : 0.f;
        for (int t = wid; t < Tb; t += nw) {
            const float* xrow = p.logits + (int64_t)b * p.sb + (int64_t)t * p.st;
            float lse = p.lse[(int64_t)b * p.T + t], H = p.hrow[(int64_t)b * p.T + t];
            for (int v = lane; v < p.V; v += 32) {
                float lp = xrow[v] - lse; float pv = __expf(lp);
                gb[(int64_t)t * p.V + v] = (pv > 0.f) ? c_kl * pv * (lp - H) : 0.f;
            }
        }
        return;
    }

    const float* al = p.alpha + base;
    const float* be = p.beta + base;
    const int bl = min(max(p.blank, 0), p.V - 1);
    for (int t = wid; t < Tb; t += nw) {
        const float* a = al + (int64_t)t * Sp;
        const float* bt = be + (int64_t)t * Sp;
        const float* e = em + (int64_t)t * Sp;
        float* grow = gb + (int64_t)t * p.V;
        // blank column: all even states
        float m = NSP_NEG_BIG, sum = 0.f;
        for (int s = 2 * lane; s < S; s += 64) {
            float v = a[s] + bt[s];
            float nm = fmaxf(m, v);
            sum = sum * __expf(m - nm) + __expf(v - nm);
            m = nm;
        }
        float M = warp_max(m);
        sum = warp_sum(sum * __expf(m - M));
        if (lane == 0) {
            float lcab = M + __logf(sum);
            grow[bl] -= c_ctc * __expf(lcab + nll - e[0]);
        }
        // label columns: head states walk their same-label chain (deterministic order)
        for (int s = 2 * lane + 1; s < S; s += 64) {
            if (head[s]) {
                float mm = a[s] + bt[s], ss = 1.f;
                for (int q = nxt[s]; q >= 0; q = nxt[q]) {
                    float v = a[q] + bt[q];
                    float nm = fmaxf(mm, v);
                    ss = ss * __expf(mm - nm) + __expf(v - nm);
                    mm = nm;
                }
                float lcab = mm + __logf(ss);
                int v = min(max(lab[s >> 1], 0), p.V - 1);
                if (v != bl) grow[v] -= c_ctc * __expf(lcab + nll - e[s]);
                else atomicAdd(&grow[v], -c_ctc * __expf(lcab + nll - e[s]));   // label == blank id (degenerate input)
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// K2 (default): warp-per-sweep lattice.  CTA = 2 warps for one utterance: warp 0 runs alpha forward, warp 1 runs beta
// backward, each lane owns K consecutive states in registers, neighbours come from __shfl_up/down_sync -- no block
// barriers on the serial chain; emissions are prefetched PF steps ahead with 128-bit loads.
// ---------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void ld_states(const float* p, float (&v)[K]) {
    if constexpr (K >= 4) {
#pragma unroll
        for (int i = 0; i < K; i += 4) {
            float4 t = *reinterpret_cast<const float4*>(p + i);
            v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
        }
    } else if constexpr (K == 2) {
        float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y;
    } else {
        v[0] = *p;
    }
}
template <int K>
__device__ __forceinline__ void st_states(float* p, const float (&v)[K]) {
    if constexpr (K >= 4) {
#pragma unroll
        for (int i = 0; i < K; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else if constexpr (K == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
        *p = v[0];
    }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int K, int NWD>
__global__ void __launch_bounds__(64 * NWD) ctc_lattice_warp_kernel(CtcParams p) {
    pdl_entry();
    // NWD warps per direction (alpha: warps [0,NWD), beta: warps [NWD,2*NWD)); each lane owns K consecutive states.
    // Inside a warp neighbours travel by shuffle; across warps through a double-buffered smem slot + a named barrier
    // per direction.  Emission rows are staged CT steps at a time into shared memory with cp.async (double buffered),
    // so the serial chain per step is: 1 LDS, 2 shuffles, one log-sum-exp, 1 STG, boundary publish, barrier.
    constexpr int CT = 8;                                     // time steps per staged chunk
    constexpr int NL = 32 * NWD;                              // lanes per direction
    extern __shared__ __align__(16) float s_em[];             // [dir][2 buffers][CT][Sp]
    __shared__ int32_t s_lab[256 + 8];
    __shared__ float s_bnd[2][2][NWD][2];                     // [dir][parity][warp][2 boundary states]
    __shared__ float s_red[NWD][2];
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_beta = warp >= NWD;
    const int wd = warp % NWD;
    const int glane = wd * 32 + lane;
    const int Sp = p.Sp;
    const int L = min(max(p.ylens[b], 0), p.Lmax);
    const int S = 2 * L + 1;
    const int Tb = min(max(p.elens[b], 0), p.T);
    const int32_t* lab = p.labels + (int64_t)b * p.Lmax;
    const int64_t base = (int64_t)b * p.T * Sp;
    const float* em = p.emit + base;
    float* gout = (is_beta ? p.beta : p.alpha) + base;
    const int dsel = is_beta ? 1 : 0;
    float* my_em = s_em + (size_t)dsel * 2 * CT * Sp;
    auto dir_barrier = [&]() {
        if (is_beta) asm volatile("bar.sync 2, %0;" :: "n"(NL) : "memory");
        else asm volatile("bar.sync 1, %0;" :: "n"(NL) : "memory");
    };

    for (int i = threadIdx.x; i < L; i += 64 * NWD) s_lab[i] = lab[i];
    __syncthreads();
    // same-label chains for the sparse gradient fix-up (consumed by ctc_fixup_kernel)
    for (int s = threadIdx.x; s < S; s += 64 * NWD) {
        int nx = -1, hd = 1;
        if (s & 1) {
            const int me = s_lab[s >> 1];
            for (int q = (s >> 1) + 1; q < L; ++q) if (s_lab[q] == me) { nx = 2 * q + 1; break; }
            for (int q = 0; q < (s >> 1); ++q) if (s_lab[q] == me) { hd = 0; break; }
        }
        p.nxt[(int64_t)b * Sp + s] = (int16_t)nx;
        p.head[(int64_t)b * Sp + s] = (int16_t)hd;
    }

    const int s0 = glane * K;
    bool valid[K], skip[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int s = s0 + k;
        valid[k] = s < S;
        skip[k] = false;
        if (valid[k] && (s & 1)) {
            if (!is_beta) skip[k] = (s >= 2) && (s_lab[s >> 1] != s_lab[(s >> 1) - 1]);
            else          skip[k] = (s + 2 < S) && (s_lab[s >> 1] != s_lab[(s >> 1) + 1]);
        }
    }
    if (Tb <= 0) {
        if (threadIdx.x == 0) { p.nll_raw[b] = (L == 0) ? 0.f : 1.0e30f; p.nll[b] = 0.f; }
        return;
    }
    const bool lane_active = s0 < S;                      // lanes beyond the path never store
    // stage chunk c (steps c*CT .. c*CT+CT-1 of this direction's sweep) into buffer c & 1
    const int vec_per_row = Sp / 4;
    auto stage = [&](int c) {
        float* dst = my_em + (size_t)(c & 1) * CT * Sp;
        for (int e = glane; e < CT * vec_per_row; e += NL) {
            const int tt = e / vec_per_row, v4 = e % vec_per_row;
            const int i = c * CT + tt;
            if (i < Tb) {
                const int t = is_beta ? (Tb - 1 - i) : i;
                cp_async16(dst + tt * Sp + v4 * 4, em + (int64_t)t * Sp + v4 * 4);
            }
        }
        cp_async_commit();
    };
    const int nchunks = (Tb + CT - 1) / CT;
    stage(0);
    if (nchunks > 1) stage(1); else cp_async_commit();

    float own[K];
#pragma unroll
    for (int k = 0; k < K; ++k) own[k] = NSP_NEG_BIG;
    const int nbw = is_beta ? wd + 1 : wd - 1;
    const bool has_nb = (nbw >= 0 && nbw < NWD);
    const int64_t gstep = is_beta ? -(int64_t)Sp : (int64_t)Sp;
    float* gptr = gout + (int64_t)(is_beta ? (Tb - 1) : 0) * Sp + s0;

    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<1>();
        dir_barrier();                                     // chunk c landed for every thread of this direction
        const float* ebuf = my_em + (size_t)(c & 1) * CT * Sp + s0;
        const int nst = min(CT, Tb - c * CT);
        for (int tt = 0; tt < nst; ++tt) {
            const int i = c * CT + tt;
            float e[K];
#pragma unroll
            for (int k = 0; k < K; ++k) e[k] = 0.f;
            if (lane_active) {                             // lanes past the row pitch must not touch the staging buffer
                if constexpr (K == 1) e[0] = ebuf[tt * Sp];
                else if constexpr (K == 2) { float2 t2 = *reinterpret_cast<const float2*>(ebuf + tt * Sp); e[0] = t2.x; e[1] = t2.y; }
                else {
#pragma unroll
                    for (int k = 0; k < K; k += 4) {
                        float4 t4 = *reinterpret_cast<const float4*>(ebuf + tt * Sp + k);
                        e[k] = t4.x; e[k + 1] = t4.y; e[k + 2] = t4.z; e[k + 3] = t4.w;
                    }
                }
            }
            float nw[K];
            if (i == 0) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int s = s0 + k;
                    const bool start = is_beta ? (s >= S - 2) : (s <= 1);
                    nw[k] = (valid[k] && start) ? e[k] : NSP_NEG_BIG;
                }
            } else {
                // the two states adjacent to this lane's block, from the previous step:
                //   alpha: n1 = state s0-1, n2 = state s0-2;   beta: n1 = state s0+K, n2 = state s0+K+1
                float n1, n2;
                float b1 = NSP_NEG_BIG, b2 = NSP_NEG_BIG;    // values owned by the neighbouring warp
                if (has_nb) { b1 = s_bnd[dsel][(i - 1) & 1][nbw][0]; b2 = s_bnd[dsel][(i - 1) & 1][nbw][1]; }
                if (!is_beta) {
                    n1 = __shfl_up_sync(0xffffffffu, own[K - 1], 1);
                    n2 = (K >= 2) ? __shfl_up_sync(0xffffffffu, own[K >= 2 ? K - 2 : 0], 1) : __shfl_up_sync(0xffffffffu, own[0], 2);
                    if (lane == 0) { n1 = b1; n2 = b2; }
                    if (K == 1 && lane == 1) n2 = b1;
                } else {
                    n1 = __shfl_down_sync(0xffffffffu, own[0], 1);
                    n2 = (K >= 2) ? __shfl_down_sync(0xffffffffu, own[K >= 2 ? 1 : 0], 1) : __shfl_down_sync(0xffffffffu, own[0], 2);
                    if (lane == 31) { n1 = b1; n2 = b2; }
                    if (K == 1 && lane == 30) n2 = b1;
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    float a1, a2;
                    if (!is_beta) {
                        a1 = (k >= 1) ? own[k >= 1 ? k - 1 : 0] : n1;
                        a2 = (k >= 2) ? own[k >= 2 ? k - 2 : 0] : ((k == 1) ? n1 : n2);
                    } else {
                        a1 = (k + 1 < K) ? own[k + 1 < K ? k + 1 : 0] : n1;
                        a2 = (k + 2 < K) ? own[k + 2 < K ? k + 2 : 0] : ((k + 2 == K) ? n1 : n2);
                    }
                    if (!skip[k]) a2 = NSP_NEG_BIG;
                    float v = lse3(own[k], a1, a2) + e[k];
                    nw[k] = valid[k] ? fmaxf(v, NSP_NEG_BIG) : NSP_NEG_BIG;
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) own[k] = nw[k];
            if (lane_active) st_states<K>(gptr, own);
            gptr += gstep;
            if constexpr (NWD > 1) {
                // publish this warp's boundary states for the neighbouring warp's next step
                if (!is_beta) {
                    if (K >= 2) { if (lane == 31) { s_bnd[0][i & 1][wd][0] = own[K - 1]; s_bnd[0][i & 1][wd][1] = own[K >= 2 ? K - 2 : 0]; } }
                    else { if (lane == 31) s_bnd[0][i & 1][wd][0] = own[0]; if (lane == 30) s_bnd[0][i & 1][wd][1] = own[0]; }
                } else {
                    if (K >= 2) { if (lane == 0) { s_bnd[1][i & 1][wd][0] = own[0]; s_bnd[1][i & 1][wd][1] = own[K >= 2 ? 1 : 0]; } }
                    else { if (lane == 0) s_bnd[1][i & 1][wd][0] = own[0]; if (lane == 1) s_bnd[1][i & 1][wd][1] = own[0]; }
                }
                dir_barrier();
            }
        }
        if constexpr (NWD == 1) dir_barrier();             // everyone done reading buffer c & 1 before it is refilled
        if (c + 2 < nchunks) stage(c + 2); else cp_async_commit();
    }
    if (!is_beta) {
        // nll = -lse(alpha_{T-1}(S-1), alpha_{T-1}(S-2)); the two states may sit in different warps
        float m = NSP_NEG_BIG;
#pragma unroll
        for (int k = 0; k < K; ++k) { const int s = s0 + k; if (s == S - 1 || s == S - 2) m = fmaxf(m, own[k]); }
        float M = warp_max(m);
        if constexpr (NWD > 1) {
            if (lane == 0) s_red[wd][0] = M;
            dir_barrier();
            M = s_red[0][0];
#pragma unroll
            for (int w = 1; w < NWD; ++w) M = fmaxf(M, s_red[w][0]);
        }
        float sm_ = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) { const int s = s0 + k; if (s == S - 1 || s == S - 2) sm_ += __expf(own[k] - M); }
        sm_ = warp_sum(sm_);
        if constexpr (NWD > 1) {
            if (lane == 0) s_red[wd][1] = sm_;
            dir_barrier();
            sm_ = 0.f;
#pragma unroll
            for (int w = 0; w < NWD; ++w) sm_ += s_red[w][1];
        }
        if (threadIdx.x == 0) {
            const float nll = -(M + __logf(sm_));
            p.nll_raw[b] = nll;
            p.nll[b] = (nll < 1.0e29f) ? nll : 0.f;
        }
    }
}

// K3: sparse gradient fix-up, one warp per (b, t) row (thousands of independent HBM read-modify-writes in flight),
// plus the scalar loss reduction in CTA 0.
__global__ void __launch_bounds__(256) ctc_fixup_kernel(CtcParams p) {
    pdl_entry();
    __shared__ float scratch[32];
    if (blockIdx.x == gridDim.x - 1) {     // dedicated last CTA: loss = (1-lsm) * sum nll / B + lsm * KL
        float a = 0.f;
        for (int i = threadIdx.x; i < p.B; i += 256) a += p.nll[i];
        a = block_sum<256>(a, scratch);
        float loss = (1.f - p.lsm) * a / (float)p.B;
        if (p.lsm > 0.f) {
            float k = 0.f;
            const int64_t n = (int64_t)p.B * p.T;
            for (int64_t i = threadIdx.x; i < n; i += 256) k += p.klrow[i];
            k = block_sum<256>(k, scratch);
            loss += p.lsm * k / n_frames_of(p.elens, p.B, p.T);
        }
        if (threadIdx.x == 0) p.loss[0] = loss;
        return;
    }
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= (int64_t)p.B * p.T) return;
    const int b = (int)(row / p.T), t = (int)(row % p.T);
    const int Tb = min(max(p.elens[b], 0), p.T);
    if (t >= Tb) return;
    const int Sp = p.Sp;
    const int L = min(max(p.ylens[b], 0), p.Lmax);
    const int S = 2 * L + 1;
    const float nll = p.nll_raw[b];
    const float c_ctc = (1.f - p.lsm) / (float)p.B;
    float* grow = p.grad + row * (int64_t)p.V;
    if (!(nll < 1.0e29f)) {
        // zero_infinity: CTC part of the gradient vanishes; only the label-smoothing KL part stays
        const float c_kl = (p.lsm > 0.f) ? p.lsm / n_frames_of(p.elens, p.B, p.T) : 0.f;
        const float* xrow = p.logits + (int64_t)b * p.sb + (int64_t)t * p.st;
        const float lse = p.lse[row], H = p.hrow[row];
        for (int v = lane; v < p.V; v += 32) {
            float lp = xrow[v] - lse; float pv = __expf(lp);
            grow[v] = (pv > 0.f) ? c_kl * pv * (lp - H) : 0.f;
        }
        return;
    }
    const float* a = p.alpha + row * (int64_t)Sp;
    const float* bt = p.beta + row * (int64_t)Sp;
    const float* e = p.emit + row * (int64_t)Sp;
    const int16_t* nxt = p.nxt + (int64_t)b * Sp;
    const int16_t* head = p.head + (int64_t)b * Sp;
    const int32_t* lab = p.labels + (int64_t)b * p.Lmax;
    const int bl = min(max(p.blank, 0), p.V - 1);
    // blank column: all even states
    float m = NSP_NEG_BIG, sum = 0.f;
    for (int s = 2 * lane; s < S; s += 64) {
        float v = a[s] + bt[s];
        float nm = fmaxf(m, v);
        sum = sum * __expf(m - nm) + __expf(v - nm);
        m = nm;
    }
    const float M = warp_max(m);
    sum = warp_sum(sum * __expf(m - M));
    if (lane == 0) grow[bl] -= c_ctc * __expf(M + __logf(sum) + nll - e[0]);
    // label columns: head states walk their same-label chain (deterministic order)
    for (int s = 2 * lane + 1; s < S; s += 64) {
        if (head[s]) {
            float mm = a[s] + bt[s], ss = 1.f;
            for (int q = nxt[s]; q >= 0; q = nxt[q]) {
                float v = a[q] + bt[q];
                float nm = fmaxf(mm, v);
                ss = ss * __expf(mm - nm) + __expf(v - nm);
                mm = nm;
            }
            const float lcab = mm + __logf(ss);
            const int v = min(max(lab[s >> 1], 0), p.V - 1);
            if (v != bl) grow[v] -= c_ctc * __expf(lcab + nll - e[s]);
            else atomicAdd(&grow[v], -c_ctc * __expf(lcab + nll - e[s]));   // label == blank id (degenerate input)
        }
    }
}

__global__ void __launch_bounds__(256) ctc_finalize_kernel(CtcParams p) {
    pdl_entry();
    __shared__ float scratch[32];
    float a = 0.f;
    for (int i = threadIdx.x; i < p.B; i += 256) a += p.nll[i];
    a = block_sum<256>(a, scratch);
    float loss = (1.f - p.lsm) * a / (float)p.B;
    if (p.lsm > 0.f) {
        float k = 0.f;
        const int64_t n = (int64_t)p.B * p.T;
        for (int64_t i = threadIdx.x; i < n; i += 256) k += p.klrow[i];
        k = block_sum<256>(k, scratch);
        loss += p.lsm * k / n_frames_of(p.elens, p.B, p.T);
    }
    if (threadIdx.x == 0) p.loss[0] = loss;
}

}  // namespace
}  // namespace nsp
#include "ctc_stream.cuh"
namespace nsp {
namespace {

template <int G, int VEC>
nsp_status launch_rows(const CtcParams& p, int nvec, cudaStream_t st) {
    const int64_t rows = (int64_t)p.B * p.T;
    const int rpb = 256 / G;
    const unsigned grid = (unsigned)ceil_div64(rows, rpb);
    const int vpt = ceil_div(nvec, G);
#define NSP_ROWS(VPT) launch_k(ctc_rows_kernel<G, VPT, VEC>, dim3(grid), dim3(256), 0, st, p)
    if (vpt <= 1) NSP_ROWS(1);
    else if (vpt <= 2) NSP_ROWS(2);
    else if (vpt <= 4) NSP_ROWS(4);
    else if (vpt <= 8) NSP_ROWS(8);
    else if (vpt <= 12) NSP_ROWS(12);
    else { set_error("ctc rows: internal dispatch error (vpt=%d)", vpt); return NSP_ERR_INVALID; }
#undef NSP_ROWS
    NSP_LAUNCH_OK();
    return NSP_OK;
}

}  // namespace
}  // namespace nsp

using namespace nsp;

// lattice row pitch: 32 K states (K = states per lane of the one-warp sweeps, ctc_stream.cuh) up to 512 path states, else a
// multiple of 16 -- a lane's K-wide vector access never crosses into the next row either way
static size_t ctc_row_pitch(int Lmax) {
    const int Smax = 2 * Lmax + 1;
    if (Smax <= 512) return (size_t)32 * (Smax <= 32 ? 1 : Smax <= 64 ? 2 : Smax <= 128 ? 4 : Smax <= 256 ? 8 : 16);
    return align_up((size_t)Smax, 16);
}

extern "C" size_t nsp_ctc_loss_workspace_bytes(int B, int T, int Lmax) {
    if (B <= 0 || T <= 0 || Lmax < 0) return 0;
    size_t Sp = ctc_row_pitch(Lmax);
    size_t bt = (size_t)B * T;
    return align_up(3 * bt * Sp * sizeof(float), 256) + 3 * align_up(bt * sizeof(float), 256) +
           align_up((size_t)B * sizeof(float), 256) + 2 * align_up((size_t)B * Sp * sizeof(int16_t), 256) +
           align_up((size_t)(B + 1) * sizeof(int32_t), 256) + 256 +
           (size_t)(2 * 160 + 4 * (size_t)B + 64) * sizeof(unsigned long long);      /* bring-up trace (NSP_CTC_DEBUG=8) */
}

extern "C" nsp_status nsp_ctc_loss_fwd_bwd(const float* logits, int64_t stride_b, int64_t stride_t,
                                           int B, int T, int V,
                                           const int32_t* labels, int Lmax,
                                           const int32_t* elens, const int32_t* ylens,
                                           int blank, float lsm_prob,
                                           float* nll, float* loss, float* grad,
                                           void* workspace, size_t workspace_bytes, void* stream) {
    NSP_CHECK_ARG(logits && elens && ylens && nll && loss && grad && workspace, "ctc_loss: null pointer");
    NSP_CHECK_ARG(labels || Lmax == 0, "ctc_loss: labels is null with Lmax=%d", Lmax);
    NSP_CHECK_ARG(B > 0 && T > 0 && V > 1, "ctc_loss: bad shape B=%d T=%d V=%d", B, T, V);
    NSP_CHECK_ARG(Lmax >= 0 && 2 * Lmax + 1 <= 4096, "ctc_loss: Lmax=%d unsupported (2L+1 <= 4096)", Lmax);
    NSP_CHECK_ARG(blank >= 0 && blank < V, "ctc_loss: blank=%d out of range", blank);
    NSP_CHECK_ARG(lsm_prob >= 0.f && lsm_prob < 1.f, "ctc_loss: lsm_prob=%f", lsm_prob);
    NSP_CHECK_ARG(workspace_bytes >= nsp_ctc_loss_workspace_bytes(B, T, Lmax), "ctc_loss: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;

    CtcParams p;
    p.logits = logits; p.sb = stride_b; p.st = stride_t; p.B = B; p.T = T; p.V = V;
    p.labels = labels; p.Lmax = Lmax; p.elens = elens; p.ylens = ylens;
    p.blank = blank; p.lsm = lsm_prob; p.nll = nll; p.loss = loss; p.grad = grad;
    p.Sp = (int)ctc_row_pitch(Lmax);
    const size_t bt = (size_t)B * T;
    char* w = (char*)workspace;
    const size_t lat = bt * p.Sp * sizeof(float);
    p.emit = (float*)w; p.alpha = (float*)(w + lat); p.beta = (float*)(w + 2 * lat);
    w += align_up(3 * lat, 256);
    p.lse = (float*)w; w += align_up(bt * sizeof(float), 256);
    p.hrow = (float*)w; w += align_up(bt * sizeof(float), 256);
    p.klrow = (float*)w; w += align_up(bt * sizeof(float), 256);
    p.nll_raw = (float*)w; w += align_up((size_t)B * sizeof(float), 256);
    p.nxt = (int16_t*)w; w += align_up((size_t)B * p.Sp * sizeof(int16_t), 256);
    p.head = (int16_t*)w; w += align_up((size_t)B * p.Sp * sizeof(int16_t), 256);
    int32_t* ready = (int32_t*)w;

    // ---- streaming path (ctc_stream.cuh): TMA-pipelined row pass with the lattice sweeps running under it ----
    {
        static const bool legacy = [] { const char* e = getenv("NSP_CTC_PATH"); return e && !strcmp(e, "legacy"); }();
        const bool al16 = (V % 4 == 0) && (stride_b % 4 == 0) && (stride_t % 4 == 0) &&
                          (((uintptr_t)logits) % 16 == 0) && (((uintptr_t)grad) % 16 == 0);
        const bool contiguous = (stride_t == V) && (stride_b == (int64_t)T * V);
        const int Smax = 2 * Lmax + 1;
        const int V4 = V / 4;
        if (!legacy && al16 && V <= 12288 && Smax <= 512 && (V4 > 320 || contiguous)) {
            CtcStream q;
            memset(&q, 0, sizeof(q));
            q.mode_warp = V4 <= 320;
            const size_t lat = (size_t)2 * 2 * CS_CT * p.Sp * sizeof(float);
            const size_t avail = (size_t)220 * 1024 - align_up(lat, 1024);
            const int64_t rows = (int64_t)bt;
            if (q.mode_warp) {
                int R = (int)(32768 / ((size_t)V * 4)) / 16 * 16;
                R = R < 16 ? 16 : (R > 256 ? 256 : R);
                const int spread = (int)(rows / num_sms()) / 16 * 16;        // small problems: use every SM
                if (spread < R) R = spread < 16 ? 16 : spread;
                q.R = R;
            } else {
                q.R = 1;
            }
            q.tile_floats = q.R * V;
            const size_t tile_bytes = (size_t)q.tile_floats * sizeof(float);
            q.stages = (int)(avail / tile_bytes);
            if (q.stages > CS_MAX_STAGES) q.stages = CS_MAX_STAGES;
            if (q.stages >= 2) {
                q.ntiles = ceil_div64(rows, q.R);
                q.K = Smax <= 32 ? 1 : Smax <= 64 ? 2 : Smax <= 128 ? 4 : Smax <= 256 ? 8 : 16;
                q.ready = ready;
                q.next_tile = ready + B;
                static const int dbg = [] { const char* e = getenv("NSP_CTC_DEBUG"); return e ? atoi(e) : 0; }();
                q.dbg = dbg;
                q.trace = (dbg & 8) ? (unsigned long long*)((char*)ready + align_up((size_t)(B + 1) * sizeof(int32_t), 256)) : nullptr;
                if (q.trace) NSP_CUDA_OK(cudaMemsetAsync(q.trace, 0, (size_t)(2 * 160 + 4 * (size_t)B + 64) * sizeof(unsigned long long), st));
                const size_t smem = (size_t)q.stages * tile_bytes + lat;
                const unsigned grid = (unsigned)(q.ntiles < num_sms() ? q.ntiles : num_sms());
                NSP_CUDA_OK(cudaMemsetAsync(ready, 0, (size_t)(B + 1) * sizeof(int32_t), st));
                static size_t attr_w = 0, attr_r = 0;          // opt-in shared memory: raise the cap only when it grows
                if (q.mode_warp) {
                    if (smem > attr_w) { NSP_CUDA_OK(cudaFuncSetAttribute(ctc_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_w = smem; }
                    launch_k(ctc_stream_kernel<true>, dim3(grid), dim3(CS_THREADS), smem, st, p, q);
                } else {
                    if (smem > attr_r) { NSP_CUDA_OK(cudaFuncSetAttribute(ctc_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_r = smem; }
                    launch_k(ctc_stream_kernel<false>, dim3(grid), dim3(CS_THREADS), smem, st, p, q);
                }
                NSP_LAUNCH_OK();
                launch_k(ctc_fixup_write_kernel, dim3((unsigned)(ceil_div64((int64_t)bt, 8) + 1)), dim3(256), 0, st, p);
                NSP_LAUNCH_OK();
                return NSP_OK;
            }
        }
    }

    // ---- K1 ----
    const bool vec4 = (V % 4 == 0) && (stride_b % 4 == 0) && (stride_t % 4 == 0) &&
                      (((uintptr_t)logits) % 16 == 0) && (((uintptr_t)grad) % 16 == 0);
    const int vec = vec4 ? 4 : 1;
    const int nvec = V / vec;
    nsp_status s;
    if (nvec <= 32 * 8) {
        s = vec4 ? launch_rows<32, 4>(p, nvec, st) : launch_rows<32, 1>(p, nvec, st);
    } else if (nvec <= 256 * 12) {
        s = vec4 ? launch_rows<256, 4>(p, nvec, st) : launch_rows<256, 1>(p, nvec, st);
    } else {
        size_t smem = (size_t)V * sizeof(float);
        if (smem > 200 * 1024) { set_error("ctc_loss: V=%d too large (max 51200)", V); return NSP_ERR_UNSUPPORTED; }
        NSP_CUDA_OK(cudaFuncSetAttribute(ctc_rows_smem_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        launch_k(ctc_rows_smem_kernel<512>, dim3((unsigned)bt), dim3(512), smem, st, p);
        NSP_LAUNCH_OK();
        s = NSP_OK;
    }
    if (s != NSP_OK) return s;

    // ---- K2 / K3 ----
    const int Smax = 2 * Lmax + 1;
    if (Smax <= 512 && Lmax <= 256) {
        // four warps per direction (one per SM sub-partition) unless the path is so short that one warp holds it
        const size_t lsm = (size_t)2 * 2 * 8 * p.Sp * sizeof(float);     // [dir][buffer][CT = 8][Sp]
        if (Smax <= 32) {
            launch_k(ctc_lattice_warp_kernel<1, 1>, dim3(B), dim3(64), lsm, st, p);
        } else {
            const int k = ceil_div(Smax, 128);
            if (lsm > 48 * 1024) {
                NSP_CUDA_OK(cudaFuncSetAttribute(ctc_lattice_warp_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm));
                NSP_CUDA_OK(cudaFuncSetAttribute(ctc_lattice_warp_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm));
            }
            if (k <= 1) launch_k(ctc_lattice_warp_kernel<1, 4>, dim3(B), dim3(256), lsm, st, p);
            else if (k <= 2) launch_k(ctc_lattice_warp_kernel<2, 4>, dim3(B), dim3(256), lsm, st, p);
            else launch_k(ctc_lattice_warp_kernel<4, 4>, dim3(B), dim3(256), lsm, st, p);
        }
        NSP_LAUNCH_OK();
        launch_k(ctc_fixup_kernel, dim3((unsigned)(ceil_div64((int64_t)bt, 8) + 1)), dim3(256), 0, st, p);
        NSP_LAUNCH_OK();
        return NSP_OK;
    }
    {   // very long label sequences: block-per-utterance lattice with the fix-up fused in
        const int S = p.Sp;
        int spt = 1, half = (int)align_up((size_t)S, 32);
        if (half > 512) { spt = 2; half = (int)align_up((size_t)ceil_div(S, 2), 32); }
        if (half > 512) { spt = 4; half = (int)align_up((size_t)ceil_div(S, 4), 32); }
        if (half > 512) { spt = 8; half = 512; }
        size_t smem = (size_t)4 * S * sizeof(float) + (size_t)2 * S * sizeof(int16_t) + 16;
        dim3 block(2 * half);
        if (smem > 48 * 1024) {
            NSP_CUDA_OK(cudaFuncSetAttribute(ctc_lattice_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            NSP_CUDA_OK(cudaFuncSetAttribute(ctc_lattice_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            NSP_CUDA_OK(cudaFuncSetAttribute(ctc_lattice_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            NSP_CUDA_OK(cudaFuncSetAttribute(ctc_lattice_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        if (spt == 1) launch_k(ctc_lattice_kernel<1>, dim3(B), dim3(block), smem, st, p, half);
        else if (spt == 2) launch_k(ctc_lattice_kernel<2>, dim3(B), dim3(block), smem, st, p, half);
        else if (spt == 4) launch_k(ctc_lattice_kernel<4>, dim3(B), dim3(block), smem, st, p, half);
        else launch_k(ctc_lattice_kernel<8>, dim3(B), dim3(block), smem, st, p, half);
        NSP_LAUNCH_OK();
    }
    launch_k(ctc_finalize_kernel, dim3(1), dim3(256), 0, st, p);
    NSP_LAUNCH_OK();
    return NSP_OK;
}
