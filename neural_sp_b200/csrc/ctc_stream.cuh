// CTC forward+backward, streaming variant (included by ctc_loss.cu; sm_100a).
//
// One persistent kernel, one CTA per SM, four roles:
//   producer thread   cp.async.bulk (TMA, non-tensor form) of logit tiles HBM -> shared memory, STAGES deep, mbarrier tx
//   16 compute warps  per row: max / sum-exp / entropy from registers, path emissions gathered from the staged row,
//                     gradient (softmax and label-smoothing parts) written IN PLACE into the tile
//   store thread      cp.async.bulk shared -> HBM of the finished tile (bulk groups; a slot is recycled when its store has
//                     finished reading shared memory)
//                     -- and publishes the tile's emission rows to the per-utterance counters (red.release.gpu): a fence on
//                     the compute warps cost 4 us per row (MEMBAR.SC + L1 invalidate, measured), here it is off their path
//   2 lattice warps   alpha (forward) and beta (backward) sweeps of the utterances this CTA owns; an utterance starts as
//                     soon as the per-utterance row counter says all of its emission rows exist -- i.e. UNDER the row pass
//                     of the following utterances; only the last utterance's sweep is exposed.  The sweep works in the
//                     log2 domain (ex2 / lg2.approx.ftz, 13 instructions per state and step; the round-1 loop had 37).
// Loads, math and stores of different tiles overlap by construction (the round-1 kernel alternated load / reduce / store
// phases inside each CTA and reached 62 % of the HBM rate).  A second small kernel then WRITES the final value of the
// <= L+1 touched columns per row (softmax part recomputed from the emission, so no read-modify-write) and reduces the loss.
//
// Tiles: ROW mode  = one row per tile, all 16 compute warps on it (1280 < V <= 12288);
//        WARP mode = R rows per tile (contiguous logits), one warp per row (V <= 1280).
#pragma once

namespace nsp {
namespace {

using tc::smem_u32;

constexpr int CS_NCW = 16;                 // compute warps
constexpr int CS_NCT = CS_NCW * 32;        // compute threads
constexpr int CS_GW = CS_NCW / 2;          // ROW mode: two groups of 8 warps work on two rows at once (their barrier and
constexpr int CS_GT = CS_GW * 32;          // shuffle latencies overlap; one 16-warp group spent 4000 cycles per row, measured)
constexpr int CS_THREADS = CS_NCT + 4 * 32;   // + producer, store/signal, alpha, beta warps (640 threads: 96 registers each)
constexpr int CS_W_PROD = CS_NCW, CS_W_STORE = CS_NCW + 1, CS_W_ALPHA = CS_NCW + 2, CS_W_BETA = CS_NCW + 3;
constexpr float CS_LOG2E = 1.4426950408889634f, CS_LN2 = 0.6931471805599453f;
constexpr int CS_MAX_STAGES = 8;
constexpr int CS_CT = 8;                   // lattice: time steps per staged emission chunk

struct CtcStream {
    int mode_warp;        // 1: warp per row, R rows per tile; 0: one row per tile
    int R;                // rows per tile
    int stages;
    int tile_floats;      // R * V
    int64_t ntiles;
    int K;                // lattice states per lane (1, 2, 4, 8, 16)
    int32_t* ready;       // [B] rows of utterance b whose emissions are in HBM (zeroed before the launch)
    int32_t* next_tile;   // dynamic tile scheduler (zeroed with `ready`): CTAs slowed by their lattice warps take fewer tiles
    float n_frames;       // sum_b min(elens[b], T)  -- computed on the device (see kernel prologue)
    int dbg;              // bring-up switches (NSP_CTC_DEBUG): 1 = no lattice sweeps, 2 = no row math (copy only), 8 = trace
    unsigned long long* trace;   // dbg & 8: [gridDim.x][2] CTA start / rows done, then [B][2 dirs][2] sweep start / end (globaltimer ns)
};

__device__ __forceinline__ void bulk_load_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cs_bar_group(int g) { asm volatile("bar.sync %0, %1;" :: "r"(1 + g), "n"(CS_GT) : "memory"); }
__device__ __forceinline__ unsigned long long cs_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int ld_acquire_s32(const int32_t* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
static __device__ __noinline__ void cs_ready_timeout(int b, int have, int want) {
    printf("ctc stream: utterance %d saw %d of %d emission rows after 2 s (block %d)\n", b, have, want, blockIdx.x);
    __trap();
}

__device__ __forceinline__ float cs_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cs_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cs_lse3(float a, float b, float c) {          // log2 domain
    const float m = fmaxf(a, fmaxf(b, c));
    return m + cs_lg2(cs_ex2(a - m) + cs_ex2(b - m) + cs_ex2(c - m));
}
__device__ __forceinline__ void cs_red_release(int32_t* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// ---- one direction of the lattice for one utterance: one warp, K consecutive states per lane in registers, neighbours by
// ---- shuffle, log2 domain.  The row pitch of emit / alpha / beta is exactly 32 K (host), states >= S carry emission
// ---- NSP_NEG_BIG (written by the row pass), so no lane or state needs a predicate inside the loop.
template <int K, bool BETA>
__device__ __forceinline__ void cs_sweep(const CtcParams& p, int b, float* my_em, const int32_t* s_lab, int lane, int S, int Tb, bool no_store) {
    constexpr int Sp = 32 * K;
    constexpr int VPR = Sp / 4;                             // 16-byte vectors per emission row
    const int64_t base = (int64_t)b * p.T * Sp;
    const float* em = p.emit + base;
    float* gout = (BETA ? p.beta : p.alpha) + base;
    const int s0 = lane * K;
    uint32_t skipm = 0, startm = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int s = s0 + k;
        if (s < S && (s & 1)) {
            if (!BETA) { if (s >= 2 && s_lab[s >> 1] != s_lab[(s >> 1) - 1]) skipm |= 1u << k; }
            else       { if (s + 2 < S && s_lab[s >> 1] != s_lab[(s >> 1) + 1]) skipm |= 1u << k; }
        }
        if (BETA ? (s >= S - 2 && s < S) : (s <= 1)) startm |= 1u << k;
    }
    auto stage = [&](int c) {
        float* dst = my_em + (size_t)(c & 1) * CS_CT * Sp;
#pragma unroll
        for (int e0 = 0; e0 < CS_CT * VPR; e0 += 32) {
            const int e = e0 + lane, tt = e / VPR, v4 = e % VPR;
            const int i = c * CS_CT + tt;
            if (i < Tb) {
                const int t = BETA ? (Tb - 1 - i) : i;
                cp_async16(dst + tt * Sp + v4 * 4, em + (int64_t)t * Sp + v4 * 4);
            }
        }
        cp_async_commit();
    };
    const int nchunks = (Tb + CS_CT - 1) / CS_CT;
    stage(0);
    if (nchunks > 1) stage(1); else cp_async_commit();
    float own[K];
#pragma unroll
    for (int k = 0; k < K; ++k) own[k] = NSP_NEG_BIG;
    const int64_t gstep = BETA ? -(int64_t)Sp : (int64_t)Sp;
    float* gptr = gout + (int64_t)(BETA ? (Tb - 1) : 0) * Sp + s0;
    auto step = [&](const float* erow, bool first) {
        float e[K], nw[K];
        ld_states<K>(erow, e);
        if (first) {
#pragma unroll
            for (int k = 0; k < K; ++k) nw[k] = ((startm >> k) & 1u) ? e[k] : NSP_NEG_BIG;
        } else {
            float n1, n2;      // alpha: states s0-1, s0-2 ; beta: states s0+K, s0+K+1 (previous step)
            if (!BETA) {
                n1 = __shfl_up_sync(0xffffffffu, own[K - 1], 1);
                n2 = (K >= 2) ? __shfl_up_sync(0xffffffffu, own[K >= 2 ? K - 2 : 0], 1) : __shfl_up_sync(0xffffffffu, own[0], 2);
                if (lane == 0) { n1 = NSP_NEG_BIG; n2 = NSP_NEG_BIG; }
                if (K == 1 && lane == 1) n2 = NSP_NEG_BIG;
            } else {
                n1 = __shfl_down_sync(0xffffffffu, own[0], 1);
                n2 = (K >= 2) ? __shfl_down_sync(0xffffffffu, own[K >= 2 ? 1 : 0], 1) : __shfl_down_sync(0xffffffffu, own[0], 2);
                if (lane == 31) { n1 = NSP_NEG_BIG; n2 = NSP_NEG_BIG; }
                if (K == 1 && lane == 30) n2 = NSP_NEG_BIG;
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float a1, a2;
                if (!BETA) {
                    a1 = (k >= 1) ? own[k >= 1 ? k - 1 : 0] : n1;
                    a2 = (k >= 2) ? own[k >= 2 ? k - 2 : 0] : ((k == 1) ? n1 : n2);
                } else {
                    a1 = (k + 1 < K) ? own[k + 1 < K ? k + 1 : 0] : n1;
                    a2 = (k + 2 < K) ? own[k + 2 < K ? k + 2 : 0] : ((k + 2 == K) ? n1 : n2);
                }
                a2 = ((skipm >> k) & 1u) ? a2 : NSP_NEG_BIG;
                nw[k] = fmaxf(cs_lse3(own[k], a1, a2) + e[k], NSP_NEG_BIG);
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) own[k] = nw[k];
        if (!no_store) st_states<K>(gptr, own);
        gptr += gstep;
    };
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<1>();
        __syncwarp();
        const float* ebuf = my_em + (size_t)(c & 1) * CS_CT * Sp + s0;
        const int nst = min(CS_CT, Tb - c * CS_CT);
        if (c > 0 && nst == CS_CT) {
#pragma unroll
            for (int tt = 0; tt < CS_CT; ++tt) step(ebuf + tt * Sp, false);
        } else {
            for (int tt = 0; tt < nst; ++tt) step(ebuf + tt * Sp, c == 0 && tt == 0);
        }
        __syncwarp();                                       // everyone done with buffer c & 1 before it is refilled
        if (c + 2 < nchunks) stage(c + 2); else cp_async_commit();
    }
    cp_async_wait<0>();
    if (!BETA) {                                            // nll = -ln2 * lse2(alpha_{T-1}(S-1), alpha_{T-1}(S-2))
        float m = NSP_NEG_BIG;
#pragma unroll
        for (int k = 0; k < K; ++k) { const int s = s0 + k; if (s == S - 1 || s == S - 2) m = fmaxf(m, own[k]); }
        const float M = warp_max(m);
        float sm_ = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) { const int s = s0 + k; if (s == S - 1 || s == S - 2) sm_ += cs_ex2(own[k] - M); }
        sm_ = warp_sum(sm_);
        if (lane == 0) {
            const float nll = -(M + cs_lg2(sm_)) * CS_LN2;
            p.nll_raw[b] = nll;
            p.nll[b] = (nll < 1.0e29f) ? nll : 0.f;
        }
    }
}

// ---- row math shared by both tile modes.  NV float4 per thread, NTHR threads on the row, `tid` this thread's rank. ----
// Emissions are stored in log2 units ((x - lse) * log2 e; states >= S get NSP_NEG_BIG) for the sweeps.
template <int NV, bool CTA>
__device__ __forceinline__ void cs_process_row(const CtcParams& p, float n_frames, float* buf, int64_t row, int b, int t,
                                               int tid, float* red /*smem [2][32] of this group*/, int& red_par, int grp) {
    constexpr int NTHR = CTA ? CS_GT : 32;
    const int V = p.V, V4 = V >> 2;
    float4* b4 = reinterpret_cast<float4*>(buf);
    const int Tb = min(max(__ldg(p.elens + b), 0), p.T);
    if (t >= Tb) {                                          // padded frame: zero gradient (the tile is stored as a whole)
        for (int i = tid; i < V4; i += NTHR) b4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid == 0) { p.klrow[row] = 0.f; p.lse[row] = 0.f; p.hrow[row] = 0.f; }
        return;
    }
    const int S = 2 * min(max(__ldg(p.ylens + b), 0), p.Lmax) + 1;
    const int32_t* lab = p.labels + (int64_t)b * p.Lmax;
    // raw logits of the path states, gathered BEFORE anything overwrites the row (CTA mode: two states per thread, S <= 512)
    float xg0 = 0.f, xg1 = 0.f;
    if constexpr (CTA) {
        if (tid < S) xg0 = buf[path_label(lab, tid, p.blank, V)];
        if (tid + CS_GT < S) xg1 = buf[path_label(lab, tid + CS_GT, p.blank, V)];
    }
    // Instruction budget (one row per ~2 us of HBM time per SM): the first version spent 22 instructions and 2 MUFU per
    // element (5300 cycles per row, measured); here ~10, with ex2.approx on pre-scaled arguments.  Logits are clamped to
    // >= -1e30 so that 0 * x stays 0 for masked (-inf) entries.
    float x[NV * 4];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int i4 = j * NTHR + tid;
        float4 v = make_float4(NSP_NEG_BIG, NSP_NEG_BIG, NSP_NEG_BIG, NSP_NEG_BIG);
        if (i4 < V4) v = b4[i4];
        x[4 * j + 0] = fmaxf(v.x, NSP_NEG_BIG); x[4 * j + 1] = fmaxf(v.y, NSP_NEG_BIG);
        x[4 * j + 2] = fmaxf(v.z, NSP_NEG_BIG); x[4 * j + 3] = fmaxf(v.w, NSP_NEG_BIG);
        m = fmaxf(fmaxf(m, fmaxf(x[4 * j], x[4 * j + 1])), fmaxf(x[4 * j + 2], x[4 * j + 3]));
    }
    m = warp_max(m);
    if constexpr (CTA) {
        float* r0 = red + red_par * 32;
        if ((tid & 31) == 0) r0[tid >> 5] = m;
        cs_bar_group(grp);
        m = r0[tid & (CS_GW - 1)];
#pragma unroll
        for (int o = CS_GW / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        red_par ^= 1;
    }
    float s = 0.f, sx = 0.f;
    const float m2 = -m * CS_LOG2E;
#pragma unroll
    for (int i = 0; i < NV * 4; ++i) {
        const float e = cs_ex2(fmaf(x[i], CS_LOG2E, m2));   // exp(x - m); 0 for the padding / masked entries
        s += e;
        sx = fmaf(e, x[i], sx);
    }
    s = warp_sum(s);
    sx = warp_sum(sx);
    if constexpr (CTA) {
        float* r0 = red + red_par * 32;
        if ((tid & 31) == 0) { r0[tid >> 5] = s; r0[16 + (tid >> 5)] = sx; }
        cs_bar_group(grp);                                   // (also: every gather above precedes every overwrite below)
        s = r0[tid & (CS_GW - 1)];
        sx = r0[16 + (tid & (CS_GW - 1))];
#pragma unroll
        for (int o = CS_GW / 2; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); sx += __shfl_xor_sync(0xffffffffu, sx, o); }
        red_par ^= 1;
    }
    const float lse = m + __logf(s);
    const float H = (p.lsm > 0.f) ? (sx / s - lse) : 0.f;   // sum_v p * lp
    float* em = p.emit + row * (int64_t)p.Sp;
    if constexpr (CTA) {
        if (tid < p.Sp) em[tid] = (tid < S) ? (fmaxf(xg0, NSP_NEG_BIG) - lse) * CS_LOG2E : NSP_NEG_BIG;
        if (tid + CS_GT < p.Sp) em[tid + CS_GT] = (tid + CS_GT < S) ? (fmaxf(xg1, NSP_NEG_BIG) - lse) * CS_LOG2E : NSP_NEG_BIG;
    } else {
        for (int st = tid; st < p.Sp; st += 32)
            em[st] = (st < S) ? (fmaxf(buf[path_label(lab, st, p.blank, V)], NSP_NEG_BIG) - lse) * CS_LOG2E : NSP_NEG_BIG;
        __syncwarp();                                        // gathers done before the row is overwritten
    }
    if (tid == 0) {
        p.lse[row] = lse;
        p.hrow[row] = H;
        p.klrow[row] = (p.lsm > 0.f) ? (H + __logf((float)(V - 1))) : 0.f;
    }
    const float c_ctc = (1.f - p.lsm) / (float)p.B;
    const float c_kl = (p.lsm > 0.f) ? p.lsm / n_frames : 0.f;
    const float c0 = c_ctc - c_kl * (lse + H);               // g = p * (c_ctc + c_kl * (x - lse - H)) = p * (c0 + c_kl * x)
    const float l2 = -lse * CS_LOG2E;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int i4 = j * NTHR + tid;
        if (i4 < V4) {
            float g[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) g[k] = cs_ex2(fmaf(x[4 * j + k], CS_LOG2E, l2)) * fmaf(c_kl, x[4 * j + k], c0);
            b4[i4] = make_float4(g[0], g[1], g[2], g[3]);
        }
    }
}

template <bool MODE_WARP>
__global__ void __launch_bounds__(CS_THREADS, 1) ctc_stream_kernel(CtcParams p, CtcStream q) {
    pdl_entry();
    extern __shared__ __align__(1024) uint8_t cs_smem[];
    __shared__ __align__(8) uint64_t full_bar[CS_MAX_STAGES], comp_bar[CS_MAX_STAGES], empty_bar[CS_MAX_STAGES];
    __shared__ float s_red[2][2 * 32];                    // per compute group
    __shared__ int32_t s_lab[2][512 + 8];
    __shared__ float s_nframes;
    __shared__ int64_t s_tile[CS_MAX_STAGES];             // tile staged in each slot (-1: no more work)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int V = p.V;
    const int64_t total_rows = (int64_t)p.B * p.T;
    const size_t tile_bytes = (size_t)q.tile_floats * sizeof(float);
    float* tiles = reinterpret_cast<float*>(cs_smem);
    float* lat_em = reinterpret_cast<float*>(cs_smem + (size_t)q.stages * tile_bytes);     // [2 dirs][2][CT][Sp]

    if (threadIdx.x == 0) {
        for (int i = 0; i < q.stages; ++i) {
            tc::mbar_init(&full_bar[i], 1);
            tc::mbar_init(&comp_bar[i], MODE_WARP ? CS_NCW : CS_GW);     // one arrival per compute warp working on the slot
            tc::mbar_init(&empty_bar[i], 1);
        }
        tc::fence_barrier_init();
        int acc = 0;
        for (int i = 0; i < p.B; ++i) acc += min(max(p.elens[i], 0), p.T);
        s_nframes = (float)acc;
    }
    __syncthreads();
    const float n_frames = s_nframes;
    if (q.trace && threadIdx.x == 0) q.trace[2 * blockIdx.x] = cs_gtime();

    auto tile_rows = [&](int64_t tile) { return (int)min((int64_t)q.R, total_rows - tile * q.R); };
    auto tile_valid = [&](int64_t tile, int nr) {           // does any row of the tile carry a real frame?
        const int64_t r0 = tile * q.R, r1 = r0 + nr - 1;
        const int b0 = (int)(r0 / p.T), b1 = (int)(r1 / p.T);
        if (b1 > b0 + 1) return true;
        const int t0 = (int)(r0 % p.T);
        if (t0 < min(max(p.elens[b0], 0), p.T)) return true;
        return b1 != b0 && min(max(p.elens[b1], 0), p.T) > 0;
    };

    if (warp < CS_NCW) {
        // ------------------------------ compute warps ------------------------------
        // ROW mode: group g = warps 8g..8g+7 takes iterations g, g+2, ... ; WARP mode: all 16 warps share every tile
        const int grp = MODE_WARP ? 0 : (warp >> 3);
        const int gtid = MODE_WARP ? (int)threadIdx.x : (int)(threadIdx.x & (CS_GT - 1));
        int red_par = 0;
        for (int it = grp;; it += (MODE_WARP ? 1 : 2)) {
            const int slot = it % q.stages;
            const uint32_t ph = (uint32_t)(it / q.stages) & 1u;
            unsigned long long* rt = (q.trace && threadIdx.x == 0 && blockIdx.x == 40 && (it >> 1) < 16) ? q.trace + 2 * 160 + 4 * (size_t)p.B + 4 * (it >> 1) : nullptr;
            if (rt) rt[0] = cs_gtime();
            tc::mbar_wait_parked(&full_bar[slot], ph);
            if (rt) rt[1] = cs_gtime();
            const int64_t tile = s_tile[slot];
            if (tile < 0) break;
            float* buf = tiles + (size_t)slot * q.tile_floats;
            const int nr = tile_rows(tile);
            if (q.dbg & 2) {
            } else if constexpr (MODE_WARP) {
                for (int r = warp; r < nr; r += CS_NCW) {
                    const int64_t row = tile * q.R + r;
                    cs_process_row<10, false>(p, n_frames, buf + (size_t)r * V, row, (int)(row / p.T), (int)(row % p.T), lane, s_red[0], red_par, 0);
                }
            } else {
                const int64_t row = tile;
                cs_process_row<12, true>(p, n_frames, buf, row, (int)(row / p.T), (int)(row % p.T), gtid, s_red[grp], red_par, grp);
            }
            if (rt) rt[2] = cs_gtime();
            tc::fence_proxy_async_smem();                    // generic-proxy writes of the tile -> visible to the bulk store
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&comp_bar[slot]);
            if (rt) rt[3] = cs_gtime();
        }
        if (q.trace && threadIdx.x == 0) q.trace[2 * blockIdx.x + 1] = cs_gtime();
    } else if (warp == CS_W_PROD) {
        // ------------------------------ producer ------------------------------
        if (lane == 0) {
            for (int it = 0;; ++it) {
                const int slot = it % q.stages;
                const uint32_t ph = (uint32_t)(it / q.stages) & 1u;
                tc::mbar_wait_parked(&empty_bar[slot], ph ^ 1u);
                const int64_t tile = atomicAdd(q.next_tile, 1);
                s_tile[slot] = tile < q.ntiles ? tile : -1;          // published by the arrive below (release, cta)
                if (tile >= q.ntiles) {                              // end of work: one sentinel per compute group
                    tc::mbar_arrive(&full_bar[slot]);
                    if (!MODE_WARP) {
                        const int slot2 = (it + 1) % q.stages;
                        tc::mbar_wait_parked(&empty_bar[slot2], ((uint32_t)((it + 1) / q.stages) & 1u) ^ 1u);
                        s_tile[slot2] = -1;
                        tc::mbar_arrive(&full_bar[slot2]);
                    }
                    break;
                }
                const int nr = tile_rows(tile);
                if (!tile_valid(tile, nr)) { tc::mbar_arrive(&full_bar[slot]); continue; }
                float* buf = tiles + (size_t)slot * q.tile_floats;
                const uint32_t bytes = (uint32_t)((size_t)nr * V * sizeof(float));
                tc::mbar_arrive_expect_tx(&full_bar[slot], bytes);
                const int64_t r0 = tile * q.R;
                const float* src = MODE_WARP ? p.logits + r0 * (int64_t)V
                                             : p.logits + (r0 / p.T) * p.sb + (r0 % p.T) * p.st;
                bulk_load_g2s(buf, src, bytes, &full_bar[slot]);
            }
        }
    } else if (warp == CS_W_STORE) {
        // ------------------------------ store ------------------------------
        if (lane == 0) {
            for (int it = 0;; ++it) {
                const int slot = it % q.stages;
                const uint32_t ph = (uint32_t)(it / q.stages) & 1u;
                tc::mbar_wait_parked(&full_bar[slot], ph);
                const int64_t tile = s_tile[slot];
                if (tile < 0) break;
                tc::mbar_wait_parked(&comp_bar[slot], ph);
                const int nr = tile_rows(tile);
                bulk_store_s2g(p.grad + tile * (int64_t)q.R * V, tiles + (size_t)slot * q.tile_floats,
                               (uint32_t)((size_t)nr * V * sizeof(float)));
                tc::bulk_commit();
                tc::bulk_wait_read<0>();                     // the tile has left shared memory (sub-microsecond): recycle the slot now --
                tc::mbar_arrive(&empty_bar[slot]);           // waiting for the NEXT store first kept only 1-2 loads in flight (measured)
                if (!(q.dbg & 2)) {                          // publish the tile's emission rows: per-utterance counters, release (gpu),
                    const int64_t r0 = tile * q.R, r1 = r0 + nr;   // cumulative over the compute warps' writes (acquired through comp_bar)
                    for (int b = (int)(r0 / p.T); (int64_t)b * p.T < r1; ++b) {
                        const int64_t lo = max(r0, (int64_t)b * p.T);
                        const int64_t hi = min(r1, (int64_t)b * p.T + min(max(p.elens[b], 0), p.T));
                        if (hi > lo) cs_red_release(q.ready + b, (int)(hi - lo));
                    }
                }
            }
            bulk_wait_all();
        }
    } else {
        // ------------------------------ lattice warps ------------------------------
        const bool is_beta = (q.dbg & 16) ? (warp == CS_W_ALPHA) : (warp == CS_W_BETA);     // dbg 16: swap the two warps
        if (q.dbg & 1) return;
        int32_t* lab_s = s_lab[is_beta ? 1 : 0];
        float* my_em = lat_em + (size_t)(is_beta ? 1 : 0) * 2 * CS_CT * p.Sp;
        for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
            const int L = min(max(p.ylens[b], 0), p.Lmax);
            const int S = 2 * L + 1;
            const int Tb = min(max(p.elens[b], 0), p.T);
            const int32_t* lab = p.labels + (int64_t)b * p.Lmax;
            __syncwarp();
            for (int i = lane; i < L; i += 32) lab_s[i] = lab[i];
            __syncwarp();
            {                                                // same-label chains for the fix-up kernel (both warps, half each)
                for (int s = lane + (is_beta ? 32 : 0); s < S; s += 64) {
                    int nx = -1, hd = 1;
                    if (s & 1) {
                        const int me = lab_s[s >> 1];
                        for (int qq = (s >> 1) + 1; qq < L; ++qq) if (lab_s[qq] == me) { nx = 2 * qq + 1; break; }
                        for (int qq = 0; qq < (s >> 1); ++qq) if (lab_s[qq] == me) { hd = 0; break; }
                    }
                    p.nxt[(int64_t)b * p.Sp + s] = (int16_t)nx;
                    p.head[(int64_t)b * p.Sp + s] = (int16_t)hd;
                }
            }
            if (Tb <= 0) {
                if (!is_beta && lane == 0) { p.nll_raw[b] = (L == 0) ? 0.f : 1.0e30f; p.nll[b] = 0.f; }
                continue;
            }
            {   // wait until every emission row of this utterance is published.  Every lane polls (same address: one
                // request per poll) so the loop is warp-uniform: a lane-0-only loop left the warp DIVERGED for the whole
                // sweep -- every shuffle then took the slow reconvergence path, 0.6-0.9 us per step instead of 0.15-0.25
                // (measured with the in-kernel trace, profiles/README.md round 2).
                const long long deadline = clock64() + 4000000000LL;
                int have;
                while ((have = ld_acquire_s32(q.ready + b)) < Tb) {
                    __nanosleep(128);
                    if (clock64() > deadline) { if (lane == 0) cs_ready_timeout(b, have, Tb); __trap(); }
                }
            }
            __syncwarp();
            unsigned long long* tr = q.trace ? q.trace + 2 * 160 + ((size_t)b * 2 + (is_beta ? 1 : 0)) * 2 : nullptr;
            if (tr && lane == 0) tr[0] = cs_gtime();
#define CS_SWEEP(KK) do { if (is_beta) cs_sweep<KK, true>(p, b, my_em, lab_s, lane, S, Tb, (q.dbg & 32) != 0); \
                          else cs_sweep<KK, false>(p, b, my_em, lab_s, lane, S, Tb, (q.dbg & 32) != 0); } while (0)
            switch (q.K) {
                case 1: CS_SWEEP(1); break;
                case 2: CS_SWEEP(2); break;
                case 4: CS_SWEEP(4); break;
                case 8: CS_SWEEP(8); break;
                default: CS_SWEEP(16); break;
            }
#undef CS_SWEEP
            if (tr && lane == 0) tr[1] = cs_gtime();
        }
    }
}

// Second kernel: final value of the touched columns (write only) + the scalar loss.  One warp per (b, t) row.
__global__ void __launch_bounds__(256) ctc_fixup_write_kernel(CtcParams p) {
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
    const float c_kl = (p.lsm > 0.f) ? p.lsm / n_frames_of(p.elens, p.B, p.T) : 0.f;
    float* grow = p.grad + row * (int64_t)p.V;
    const float H = p.hrow[row];
    if (!(nll < 1.0e29f)) {
        // zero_infinity: CTC part of the gradient vanishes; only the label-smoothing KL part stays
        const float* xrow = p.logits + (int64_t)b * p.sb + (int64_t)t * p.st;
        const float lse = p.lse[row];
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
    auto soft = [&](float lp) { const float pv = __expf(lp); return (pv > 0.f) ? pv * (c_ctc + c_kl * (lp - H)) : 0.f; };
    // alpha / beta / emissions are in log2 units (ctc_stream_kernel); nll is natural
    const float nll2 = nll * CS_LOG2E;
    // blank column: all even states
    float m = NSP_NEG_BIG, sum = 0.f;
    for (int s = 2 * lane; s < S; s += 64) {
        const float v = a[s] + bt[s];
        const float nm = fmaxf(m, v);
        sum = sum * cs_ex2(m - nm) + cs_ex2(v - nm);
        m = nm;
    }
    const float M = warp_max(m);
    sum = warp_sum(sum * cs_ex2(m - M));
    // label columns: head states walk their same-label chain (deterministic order); a label equal to the blank id
    // (degenerate input) is folded into the blank column's value
    float extra = 0.f;
    for (int s = 2 * lane + 1; s < S; s += 64) {
        if (head[s]) {
            float mm = a[s] + bt[s], ss = 1.f;
            int guard = 0;
            for (int q = nxt[s]; q >= 0 && q < Sp && guard < L; q = nxt[q], ++guard) {
                const float v = a[q] + bt[q];
                const float nm = fmaxf(mm, v);
                ss = ss * cs_ex2(mm - nm) + cs_ex2(v - nm);
                mm = nm;
            }
            const float lcab2 = mm + cs_lg2(ss);
            const int v = min(max(lab[s >> 1], 0), p.V - 1);
            const float occ = c_ctc * cs_ex2(lcab2 + nll2 - e[s]);
            if (v != bl) grow[v] = soft(e[s] * CS_LN2) - occ;
            else extra += occ;
        }
    }
    extra = warp_sum(extra);
    if (lane == 0) grow[bl] = soft(e[0] * CS_LN2) - c_ctc * cs_ex2(M + cs_lg2(sum) + nll2 - e[0]) - extra;
}

}  // namespace
}  // namespace nsp
