// CTC forward+backward, streaming variant (included by ctc_loss.cu; sm_100a).
//
// One persistent kernel, one CTA per SM, four roles:
//   producer thread   cp.async.bulk (TMA, non-tensor form) of logit tiles HBM -> shared memory, STAGES deep, mbarrier tx
//   16 compute warps  per row: max / sum-exp / entropy from registers, path emissions gathered from the staged row,
//                     gradient (softmax and label-smoothing parts) written IN PLACE into the tile
//   store thread      cp.async.bulk shared -> HBM of the finished tile (bulk groups; a slot is recycled when its store has
//                     finished reading shared memory)
//   2 lattice warps   alpha (forward) and beta (backward) sweeps of the utterances this CTA owns; an utterance starts as
//                     soon as the per-utterance row counter says all of its emission rows exist -- i.e. UNDER the row pass
//                     of the following utterances; only the last utterance's sweep is exposed.
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
constexpr int CS_THREADS = CS_NCT + 4 * 32;   // + producer warp, store warp, alpha warp, beta warp
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
    float n_frames;       // sum_b min(elens[b], T)  -- computed on the device (see kernel prologue)
    int dbg;              // bring-up switches (NSP_CTC_DEBUG): 1 = no lattice sweeps, 2 = no row math (copy only)
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
__device__ __forceinline__ void cs_bar_compute() { asm volatile("bar.sync 1, %0;" :: "n"(CS_NCT) : "memory"); }
__device__ __forceinline__ int ld_acquire_s32(const int32_t* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
static __device__ __noinline__ void cs_ready_timeout(int b, int have, int want) {
    printf("ctc stream: utterance %d saw %d of %d emission rows after 2 s (block %d)\n", b, have, want, blockIdx.x);
    __trap();
}

// ---- one direction of the lattice for one utterance, one warp, K consecutive states per lane (registers + shuffles) ----
template <int K>
__device__ __forceinline__ void cs_lattice_sweep(const CtcParams& p, int b, bool is_beta, float* my_em, const int32_t* s_lab,
                                                  int lane, int L, int S, int Tb) {
    const int Sp = p.Sp;
    const int64_t base = (int64_t)b * p.T * Sp;
    const float* em = p.emit + base;
    float* gout = (is_beta ? p.beta : p.alpha) + base;
    const int s0 = lane * K;
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
    const bool lane_active = s0 < S;
    const int vec_per_row = Sp / 4;
    auto stage = [&](int c) {
        float* dst = my_em + (size_t)(c & 1) * CS_CT * Sp;
        for (int e = lane; e < CS_CT * vec_per_row; e += 32) {
            const int tt = e / vec_per_row, v4 = e % vec_per_row;
            const int i = c * CS_CT + tt;
            if (i < Tb) {
                const int t = is_beta ? (Tb - 1 - i) : i;
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
    const int64_t gstep = is_beta ? -(int64_t)Sp : (int64_t)Sp;
    float* gptr = gout + (int64_t)(is_beta ? (Tb - 1) : 0) * Sp + s0;
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<1>();
        __syncwarp();
        const float* ebuf = my_em + (size_t)(c & 1) * CS_CT * Sp + s0;
        const int nst = min(CS_CT, Tb - c * CS_CT);
        for (int tt = 0; tt < nst; ++tt) {
            const int i = c * CS_CT + tt;
            float e[K];
#pragma unroll
            for (int k = 0; k < K; ++k) e[k] = 0.f;
            if (lane_active) ld_states<K>(ebuf + tt * Sp, e);
            float nw[K];
            if (i == 0) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int s = s0 + k;
                    const bool start = is_beta ? (s >= S - 2) : (s <= 1);
                    nw[k] = (valid[k] && start) ? e[k] : NSP_NEG_BIG;
                }
            } else {
                float n1, n2;      // alpha: states s0-1, s0-2 ; beta: states s0+K, s0+K+1 (previous step)
                if (!is_beta) {
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
                    if (!is_beta) {
                        a1 = (k >= 1) ? own[k >= 1 ? k - 1 : 0] : n1;
                        a2 = (k >= 2) ? own[k >= 2 ? k - 2 : 0] : ((k == 1) ? n1 : n2);
                    } else {
                        a1 = (k + 1 < K) ? own[k + 1 < K ? k + 1 : 0] : n1;
                        a2 = (k + 2 < K) ? own[k + 2 < K ? k + 2 : 0] : ((k + 2 == K) ? n1 : n2);
                    }
                    if (!skip[k]) a2 = NSP_NEG_BIG;
                    const float v = lse3(own[k], a1, a2) + e[k];
                    nw[k] = valid[k] ? fmaxf(v, NSP_NEG_BIG) : NSP_NEG_BIG;
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) own[k] = nw[k];
            if (lane_active) st_states<K>(gptr, own);
            gptr += gstep;
        }
        __syncwarp();                                       // everyone done with buffer c & 1 before it is refilled
        if (c + 2 < nchunks) stage(c + 2); else cp_async_commit();
    }
    cp_async_wait<0>();
    if (!is_beta) {                                         // nll = -lse(alpha_{T-1}(S-1), alpha_{T-1}(S-2))
        float m = NSP_NEG_BIG;
#pragma unroll
        for (int k = 0; k < K; ++k) { const int s = s0 + k; if (s == S - 1 || s == S - 2) m = fmaxf(m, own[k]); }
        const float M = warp_max(m);
        float sm_ = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) { const int s = s0 + k; if (s == S - 1 || s == S - 2) sm_ += __expf(own[k] - M); }
        sm_ = warp_sum(sm_);
        if (lane == 0) {
            const float nll = -(M + __logf(sm_));
            p.nll_raw[b] = nll;
            p.nll[b] = (nll < 1.0e29f) ? nll : 0.f;
        }
    }
}

// ---- row math shared by both tile modes.  NV float4 per thread, `nthr` threads on the row, `tid` this thread's rank. ----
struct RowOut { float lse, H; };

template <int NV, bool CTA>
__device__ __forceinline__ void cs_process_row(const CtcParams& p, const CtcStream& q, float* buf, int64_t row, int b, int t,
                                               int tid, float* red /*smem [2][2][16]*/, int& red_par) {
    constexpr int NTHR = CTA ? CS_NCT : 32;
    const int V = p.V, V4 = V >> 2;
    float4* b4 = reinterpret_cast<float4*>(buf);
    const int Tb = min(max(p.elens[b], 0), p.T);
    if (t >= Tb) {                                          // padded frame: zero gradient (the tile is stored as a whole)
        for (int i = tid; i < V4; i += NTHR) b4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid == 0) { p.klrow[row] = 0.f; p.lse[row] = 0.f; p.hrow[row] = 0.f; }
        return;
    }
    float4 xv[NV];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int i4 = j * NTHR + tid;
        if (i4 < V4) {
            xv[j] = b4[i4];
            m = fmaxf(fmaxf(m, fmaxf(xv[j].x, xv[j].y)), fmaxf(xv[j].z, xv[j].w));
        } else {
            xv[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
    }
    m = warp_max(m);
    if constexpr (CTA) {
        float* r0 = red + red_par * 32;
        if ((tid & 31) == 0) r0[tid >> 5] = m;
        cs_bar_compute();
        m = r0[tid & 15];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        red_par ^= 1;
    }
    float s = 0.f, sx = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const float x[4] = {xv[j].x, xv[j].y, xv[j].z, xv[j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float e = __expf(x[k] - m);              // exp(-inf) = 0 for the tail / -inf logits
            s += e;
            sx += (e > 0.f) ? e * x[k] : 0.f;
        }
    }
    s = warp_sum(s);
    sx = warp_sum(sx);
    if constexpr (CTA) {
        float* r0 = red + red_par * 32;
        if ((tid & 31) == 0) { r0[tid >> 5] = s; r0[16 + (tid >> 5)] = sx; }
        cs_bar_compute();
        s = r0[tid & 15];
        sx = r0[16 + (tid & 15)];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); sx += __shfl_xor_sync(0xffffffffu, sx, o); }
        red_par ^= 1;
    }
    const float lse = m + __logf(s);
    const float H = (p.lsm > 0.f) ? (sx / s - lse) : 0.f;   // sum_v p * lp
    // path emissions from the staged row (still the logits)
    const int S = 2 * min(max(p.ylens[b], 0), p.Lmax) + 1;
    const int32_t* lab = p.labels + (int64_t)b * p.Lmax;
    float* em = p.emit + row * (int64_t)p.Sp;
    for (int st = tid; st < S; st += NTHR) em[st] = buf[path_label(lab, st, p.blank, V)] - lse;
    if (tid == 0) {
        p.lse[row] = lse;
        p.hrow[row] = H;
        p.klrow[row] = (p.lsm > 0.f) ? (H + __logf((float)(V - 1))) : 0.f;
    }
    if constexpr (CTA) cs_bar_compute(); else __syncwarp();     // gathers done before the row is overwritten
    const float c_ctc = (1.f - p.lsm) / (float)p.B;
    const float c_kl = (p.lsm > 0.f) ? p.lsm / q.n_frames : 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int i4 = j * NTHR + tid;
        if (i4 < V4) {
            const float x[4] = {xv[j].x, xv[j].y, xv[j].z, xv[j].w};
            float g[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float lp = x[k] - lse;
                const float pv = __expf(lp);
                g[k] = (pv > 0.f) ? pv * (c_ctc + c_kl * (lp - H)) : 0.f;
            }
            b4[i4] = make_float4(g[0], g[1], g[2], g[3]);
        }
    }
    if (tid == 0) {                                         // this row's emissions are published
        __threadfence();
        atomicAdd(q.ready + b, 1);
    }
}

template <bool MODE_WARP>
__global__ void __launch_bounds__(CS_THREADS, 1) ctc_stream_kernel(CtcParams p, CtcStream q) {
    extern __shared__ __align__(1024) uint8_t cs_smem[];
    __shared__ __align__(8) uint64_t full_bar[CS_MAX_STAGES], comp_bar[CS_MAX_STAGES], empty_bar[CS_MAX_STAGES];
    __shared__ float s_red[2 * 32];
    __shared__ int32_t s_lab[2][512 + 8];
    __shared__ float s_nframes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int V = p.V;
    const int64_t total_rows = (int64_t)p.B * p.T;
    const size_t tile_bytes = (size_t)q.tile_floats * sizeof(float);
    float* tiles = reinterpret_cast<float*>(cs_smem);
    float* lat_em = reinterpret_cast<float*>(cs_smem + (size_t)q.stages * tile_bytes);     // [2 dirs][2][CT][Sp]

    if (threadIdx.x == 0) {
        for (int i = 0; i < q.stages; ++i) {
            tc::mbar_init(&full_bar[i], 1);
            tc::mbar_init(&comp_bar[i], CS_NCT);
            tc::mbar_init(&empty_bar[i], 1);
        }
        tc::fence_barrier_init();
        int acc = 0;
        for (int i = 0; i < p.B; ++i) acc += min(max(p.elens[i], 0), p.T);
        s_nframes = (float)acc;
    }
    __syncthreads();
    q.n_frames = s_nframes;

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
        const int ctid = threadIdx.x;
        int red_par = 0;
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < q.ntiles; tile += gridDim.x, ++it) {
            const int slot = it % q.stages;
            const uint32_t ph = (uint32_t)(it / q.stages) & 1u;
            tc::mbar_wait(&full_bar[slot], ph);
            float* buf = tiles + (size_t)slot * q.tile_floats;
            const int nr = tile_rows(tile);
            if (q.dbg & 2) {
            } else if constexpr (MODE_WARP) {
                for (int r = warp; r < nr; r += CS_NCW) {
                    const int64_t row = tile * q.R + r;
                    cs_process_row<10, false>(p, q, buf + (size_t)r * V, row, (int)(row / p.T), (int)(row % p.T), lane, s_red, red_par);
                }
            } else {
                const int64_t row = tile;
                cs_process_row<6, true>(p, q, buf, row, (int)(row / p.T), (int)(row % p.T), ctid, s_red, red_par);
            }
            tc::fence_proxy_async_smem();                    // generic-proxy writes of the tile -> visible to the bulk store
            tc::mbar_arrive(&comp_bar[slot]);
        }
    } else if (warp == CS_NCW) {
        // ------------------------------ producer ------------------------------
        if (lane == 0) {
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < q.ntiles; tile += gridDim.x, ++it) {
                const int slot = it % q.stages;
                const uint32_t ph = (uint32_t)(it / q.stages) & 1u;
                tc::mbar_wait(&empty_bar[slot], ph ^ 1u);
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
    } else if (warp == CS_NCW + 1) {
        // ------------------------------ store ------------------------------
        if (lane == 0) {
            int it = 0, prev_slot = -1;
            for (int64_t tile = blockIdx.x; tile < q.ntiles; tile += gridDim.x, ++it) {
                const int slot = it % q.stages;
                const uint32_t ph = (uint32_t)(it / q.stages) & 1u;
                tc::mbar_wait(&comp_bar[slot], ph);
                const int nr = tile_rows(tile);
                bulk_store_s2g(p.grad + tile * (int64_t)q.R * V, tiles + (size_t)slot * q.tile_floats,
                               (uint32_t)((size_t)nr * V * sizeof(float)));
                tc::bulk_commit();
                if (prev_slot >= 0) {
                    tc::bulk_wait_read<1>();                 // the previous tile's store has left shared memory
                    tc::mbar_arrive(&empty_bar[prev_slot]);
                }
                prev_slot = slot;
            }
            bulk_wait_all();
        }
    } else {
        // ------------------------------ lattice warps ------------------------------
        const bool is_beta = warp == CS_NCW + 3;
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
            if (!is_beta) {                                  // same-label chains for the fix-up kernel
                for (int s = lane; s < S; s += 32) {
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
            if (lane == 0) {                                 // wait until every emission row of this utterance is published
                const long long deadline = clock64() + 4000000000LL;
                int have;
                while ((have = ld_acquire_s32(q.ready + b)) < Tb) {
                    __nanosleep(256);
                    if (clock64() > deadline) cs_ready_timeout(b, have, Tb);
                }
            }
            __syncwarp();
            switch (q.K) {
                case 1: cs_lattice_sweep<1>(p, b, is_beta, my_em, lab_s, lane, L, S, Tb); break;
                case 2: cs_lattice_sweep<2>(p, b, is_beta, my_em, lab_s, lane, L, S, Tb); break;
                case 4: cs_lattice_sweep<4>(p, b, is_beta, my_em, lab_s, lane, L, S, Tb); break;
                case 8: cs_lattice_sweep<8>(p, b, is_beta, my_em, lab_s, lane, L, S, Tb); break;
                default: cs_lattice_sweep<16>(p, b, is_beta, my_em, lab_s, lane, L, S, Tb); break;
            }
        }
    }
}

// Second kernel: final value of the touched columns (write only) + the scalar loss.  One warp per (b, t) row.
__global__ void __launch_bounds__(256) ctc_fixup_write_kernel(CtcParams p) {
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
    // label columns: head states walk their same-label chain (deterministic order); a label equal to the blank id
    // (degenerate input) is folded into the blank column's value
    float extra = 0.f;
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
            const float occ = c_ctc * __expf(lcab + nll - e[s]);
            if (v != bl) grow[v] = soft(e[s]) - occ;
            else extra += occ;
        }
    }
    extra = warp_sum(extra);
    if (lane == 0) grow[bl] = soft(e[0]) - c_ctc * __expf(M + __logf(sum) + nll - e[0]) - extra;
}

}  // namespace
}  // namespace nsp
