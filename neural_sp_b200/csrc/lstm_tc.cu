// (Bi)LSTM layer recurrence on the tensor cores: the same recurrence as lstm.cu (fp32 SIMT, kept for fp32 mode and the
// shapes not covered here) with the per-step product  h_{t-1} W_hh^T  issued as tcgen05.mma, bf16 operands / fp32 accumulation
// in TMEM, cell state and gate math in fp32.  This is the bf16-mode path of
//   Padding.forward  encoders/rnn.py:534-546  (pack_padded_sequence -> nn.LSTM -> pad_packed_sequence), one layer,
// and of autograd through it (BPTT).  The reference's AMP runs cuDNN's LSTM with fp16 operands; same contract.
//
// Decomposition (both kernels): all CTAs of a direction are co-resident (cooperative launch); a CTA owns UPC hidden units and
// keeps its slice of W_hh in shared memory as a K-major SWIZZLE_128B bf16 B-operand for the whole sequence.  Per time step
//   producer thread : waits for the direction's step counter (ld.acquire.gpu), then TMA-loads the step's A operand
//                     ([B rows] x 64-column boxes of the bf16 exchange buffer, L2 hits) through a ring of shared-memory stages;
//   MMA thread      : M = 64 (B <= 64) or 128 rows = batch, N = the CTA's gate columns, K = 16 per instruction, accumulates the
//                     step's [B, N] tile in TMEM; operand rows >= B are whatever lies behind the ring (never read back);
//   4 epilogue warps: TMEM -> registers -> shared memory -> one (batch, unit) item per thread: cell update (forward) or the
//                     reduction result dh_rec (backward), global stores, then thread 0 publishes the step
//                     (red.release.gpu on the counter).
// Forward :  A = h_{t-1} [B, H],   B = W_hh rows of the CTA's 4 x 8 gates  [32, H]   -> pre-activations [B, 32]
// Backward:  A = dG_t    [B, 4H],  B = W_hh^T rows of the CTA's 8 (16) units [8, 4H]   -> dh_rec [B, 8]
// Exchange buffers (bf16, double-buffered per direction) are written with generic stores by every CTA and read by TMA in every
// CTA: writer and reader both issue fence.proxy.async around the release / acquire pair.
#include <cstdlib>
#include "tc_common.cuh"

namespace nsp {
namespace {

constexpr int LT_THREADS = 192;       // warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer (+ TMEM allocation)
constexpr int LT_EPI = 128;
constexpr int FWD_UPC = 8;            // forward: 8 units = 32 gate columns per CTA
constexpr int BWD_UPC_DEFAULT = 16;   // backward: 8 units per CTA when H / 8 CTAs are co-resident, else 16 (NSP_LSTM_TC_BWD_UPC overrides)
constexpr int LT_MAXST = 8;
constexpr int LT_TMEM_COLS = 128;      // up to 4 accumulators of 32 columns: consecutive MMAs of a step go to different accumulators
                                      // (a chain into ONE accumulator runs at the MMA latency, ~70 cycles per instruction at N <= 32)
constexpr int LT_PAD = 16 * 1024;     // operand rows >= B of the last chunk read this far past the ring / W (M = 128: 16 groups x 1 KiB)

struct LtFwd {
    CUtensorMap amap;                 // bf16 [ndir*2*B, H], box [B, 64]
    const float* gx; const float* whh; const int32_t* lens; float* y;
    __nv_bfloat16* abuf;              // [ndir][2][B][H]
    unsigned int* bar;                // [ndir]
    float* acts; float* cprev; float* hprev;
    const float* h0; const float* c0; float* hN; float* cN;
    int B, T, H, ndir, dir0, ch, nst, dbg;
};

struct LtBwd {
    CUtensorMap amap;                 // bf16 [ndir*2*B, 4H], box [B, 64]
    const float* dy; const float* acts; const float* cprev; const float* whh; const int32_t* lens; float* dg;
    __nv_bfloat16* abuf;              // [ndir][2][B][4H]
    unsigned int* bar;
    const float* dhN; const float* dcN; float* dh0; float* dc0;
    int B, T, H, ndir, dir0, ch, nst, dbg;
};

__device__ __forceinline__ void lt_bar(int n) { asm volatile("bar.sync 1, %0;" :: "r"(n) : "memory"); }
__device__ __forceinline__ unsigned lt_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lt_red_release(unsigned* p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(p) : "memory");
}
__device__ __forceinline__ void lt_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Publishing a step: every epilogue thread stores its exchange values, issues fence.proxy.async (the readers are TMA loads:
// async proxy) and joins the epilogue's named barrier; thread 0 then releases the counter at gpu scope (cumulative over what
// the barrier made visible to it).  Everything that only the NEXT kernel reads (y, saved activations, dG fp32) is stored
// after the release, off the critical path of the other CTAs.
__device__ __forceinline__ void lt_publish(unsigned* bar) { lt_red_release(bar); }
__device__ __forceinline__ void lt_grid_wait(const unsigned* bar, unsigned target, int step) {
    const long long deadline = clock64() + 4000000000LL;            // ~2 s: a lost CTA becomes an error, not a hang
    while (lt_ld_acquire(bar) < target) {
        if (clock64() > deadline) { printf("lstm_tc: step counter timeout (block %d, step %d)\n", blockIdx.x, step); __trap(); }
    }
    lt_fence_proxy_async();
}
__device__ __forceinline__ float lt_sigmoid(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float lt_tanh(float x) {
    const float e = __expf(-2.f * fabsf(x));
    const float r = (1.f - e) / (1.f + e);
    return copysignf(r, x);
}

// byte offset of element (row n, column k) of a K-major SWIZZLE_128B bf16 operand whose 64-column chunks are `chunk` bytes apart
__device__ __forceinline__ uint32_t lt_sw_off(int n, int k, int chunk) {
    return (uint32_t)((k >> 6) * chunk + n * 128 + ((((k & 63) >> 3) ^ (n & 7)) << 4) + (k & 7) * 2);
}

struct LtSmem {
    uint8_t* ring; uint8_t* w; float* pre; uint64_t* full; uint64_t* empty; uint64_t* tfull; uint32_t* holder;
};
__device__ __forceinline__ LtSmem lt_carve(uint8_t* raw, int nst, int stage_bytes, int w_bytes, int pre_floats) {
    LtSmem s;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    s.ring = base;
    s.w = base + (size_t)nst * stage_bytes;
    uint8_t* q = s.w + w_bytes + LT_PAD;
    s.pre = reinterpret_cast<float*>(q);
    q += ((size_t)pre_floats * 4 + 15) & ~(size_t)15;
    s.full = reinterpret_cast<uint64_t*>(q);
    s.empty = s.full + LT_MAXST;
    s.tfull = s.empty + LT_MAXST;
    s.holder = reinterpret_cast<uint32_t*>(s.tfull + 1);
    return s;
}

// producer + MMA loops shared by both kernels: `steps` products of [B, K] x [N, K]^T, A of step i at rows row0(i) of the map
template <int M, int N>
__device__ __forceinline__ void lt_producer(const CUtensorMap* amap, const LtSmem& sm, const unsigned* bar, unsigned nctas,
                                            int steps, int B, int dirslot0, int kchunks, int ch, int nst, int chunkA, int dbg) {
    int stage = 0; uint32_t phase = 0;
    const int nit = kchunks / ch;
    const int stageA = ch * chunkA;
    for (int i = 0; i < steps; ++i) {
        lt_grid_wait(bar, (unsigned)(i + 1) * nctas, i);
        const int row = (dirslot0 + (i & 1)) * B;
        for (int it = 0; it < nit; ++it) {
            tc::mbar_wait(&sm.empty[stage], phase ^ 1);
            if (dbg & 2) { tc::mbar_arrive(&sm.full[stage]); }          // timing experiment: no operand traffic
            else {
                tc::mbar_arrive_expect_tx(&sm.full[stage], (uint32_t)(ch * B * 128));
                for (int c = 0; c < ch; ++c)
                    tc::tma_load_2d(sm.ring + (size_t)stage * stageA + (size_t)c * chunkA, amap, &sm.full[stage], (it * ch + c) * 64, row);
            }
            if (++stage == nst) { stage = 0; phase ^= 1; }
        }
    }
}
// The issuing thread's own instruction stream bounds these small MMAs (N <= 32: ~8 cycles of tensor-pipe work each): measured
// 77 cycles per instruction with descriptors rebuilt per MMA, 160 with a runtime modulo in the loop (profiles/README.md,
// round 2).  So: descriptors are built once and advanced by adds, the four K = 16 steps of a 64-column chunk are unrolled with
// compile-time offsets, and each of the four goes to its OWN accumulator (TMEM columns 0 / 32 / 64 / 96) so that consecutive
// instructions do not wait for each other's result; the epilogue adds the four partial tiles.
template <int M, int N>
__device__ __forceinline__ void lt_mma(const LtSmem& sm, uint32_t tmem, int steps, int kchunks, int ch, int nst, int chunkA, int dbg) {
    constexpr uint32_t idesc = tc::make_idesc(1u, M, N);
    int stage = 0; uint32_t phase = 0;
    const int nit = kchunks / ch;
    const uint64_t ad0 = tc::make_smem_desc_sw128(tc::smem_u32(sm.ring)), bd0 = tc::make_smem_desc_sw128(tc::smem_u32(sm.w));
    const uint64_t a_chunk = (uint64_t)(chunkA >> 4), a_stage = (uint64_t)((ch * chunkA) >> 4);
    constexpr uint64_t b_chunk = (uint64_t)((N * 128) >> 4);
    if (dbg & 1) ch = 0;                                           // timing experiment without the MMAs
    for (int i = 0; i < steps; ++i) {
        uint64_t bd = bd0;
        uint32_t accum = 0;
        for (int it = 0; it < nit; ++it) {
            tc::mbar_wait(&sm.full[stage], phase);
            tc::tc_fence_after();
            uint64_t ad = ad0 + (uint64_t)stage * a_stage;
            for (int c = 0; c < ch; ++c) {
#pragma unroll
                for (int k = 0; k < 4; ++k) tc::umma_f16(tmem + (uint32_t)(k * 32), ad + 2 * k, bd + 2 * k, idesc, accum);
                accum = 1;
                ad += a_chunk; bd += b_chunk;
            }
            tc::umma_commit(&sm.empty[stage]);
            if (++stage == nst) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit(sm.tfull);
    }
}

// accumulator tile [B, NC] TMEM -> pre[b * (NC + 1) + c]   (epilogue warps; M = 64 keeps rows 16w..16w+15 in lanes 0-15 of warp w)
template <int M, int NC>
__device__ __forceinline__ void lt_tmem_to_smem(uint32_t tmem, float* pre, int B, int warp, int lane) {
    constexpr int RPW = M == 128 ? 32 : 16;
    if (warp * RPW < B) {
        float acc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {                             // the step's partial sums, one per accumulator
            uint32_t r[32];
            if constexpr (NC > 16) tc::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * 32), r);
            else tc::tmem_ld_32x16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * 32), r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[c] += __uint_as_float(r[c]);
        }
        const int row = warp * RPW + lane;
        if (lane < RPW && row < B) {
#pragma unroll
            for (int c = 0; c < NC; ++c) pre[row * (NC + 1) + c] = acc[c];
        }
    }
    tc::tc_fence_before();
}

template <int M>
__global__ void __launch_bounds__(LT_THREADS, 1) lstm_tc_fwd_kernel(const __grid_constant__ LtFwd p) {
    pdl_entry();
    constexpr int N = 4 * FWD_UPC;
    constexpr int ITEMS = M * FWD_UPC / LT_EPI;
    extern __shared__ __align__(1024) uint8_t lt_smem_raw[];
    const int H = p.H, B = p.B;
    const int chunkA = ((B + 7) & ~7) * 128;
    const int kchunks = H / 64;
    const LtSmem sm = lt_carve(lt_smem_raw, p.nst, p.ch * chunkA, N * H * 2, B * (N + 1));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int per_dir = H / FWD_UPC;
    const int dir = p.dir0 + blockIdx.x / per_dir;
    const int j0 = (blockIdx.x % per_dir) * FWD_UPC;
    const int G = p.ndir * 4 * H;
    unsigned* bar = p.bar + dir;

    if (warp == 4 && lane == 0) tc::tma_prefetch_desc(&p.amap);
    if (warp == 5 && lane == 0) {
        for (int i = 0; i < LT_MAXST; ++i) { tc::mbar_init(&sm.full[i], 1); tc::mbar_init(&sm.empty[i], 1); }
        tc::mbar_init(sm.tfull, 1);
        tc::fence_barrier_init();
    }
    if (warp == 5) tc::tmem_alloc<LT_TMEM_COLS>(sm.holder);
    {   // resident B operand: row n = gate * 8 + u  <-  W_hh[dir][gate * H + j0 + u][:], bf16
        const float* wg = p.whh + (size_t)dir * 4 * H * H;
        const int k8n = H / 8;
        for (int e = tid; e < N * k8n; e += LT_THREADS) {
            const int n = e / k8n, k = (e % k8n) * 8;
            const float* src = wg + (size_t)((n / FWD_UPC) * H + j0 + (n % FWD_UPC)) * H + k;
            const float4 a = __ldg(reinterpret_cast<const float4*>(src));
            const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
            tc::sts128(tc::smem_u32(sm.w) + lt_sw_off(n, k, N * 128), tc::pack_bf16x2(a.x, a.y), tc::pack_bf16x2(a.z, a.w),
                       tc::pack_bf16x2(b.x, b.y), tc::pack_bf16x2(b.z, b.w));
        }
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *sm.holder;

    if (warp == 4) {
        if (tc::elect_one())
            lt_producer<M, N>(&p.amap, sm, bar, (unsigned)per_dir, p.T, B, dir * 2, kchunks, p.ch, p.nst, chunkA, p.dbg);
    } else if (warp == 5) {
        if (tc::elect_one()) lt_mma<M, N>(sm, tmem, p.T, kchunks, p.ch, p.nst, chunkA, p.dbg);
    } else {
        // item = (batch b, unit u): consecutive threads -> consecutive units of one utterance
        float c_st[ITEMS], h_st[ITEMS];
        int len[ITEMS];
        __nv_bfloat16* hb = p.abuf + (size_t)dir * 2 * B * H;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int idx = tid + LT_EPI * k, b = idx / FWD_UPC, u = idx % FWD_UPC;
            c_st[k] = 0.f; h_st[k] = 0.f; len[k] = 0;
            if (b < B) {
                len[k] = min(max(p.lens[b], 0), p.T);
                const size_t si = ((size_t)dir * B + b) * H + j0 + u;
                if (p.c0) c_st[k] = __ldg(p.c0 + si);
                if (p.h0) h_st[k] = __ldg(p.h0 + si);
                hb[(size_t)b * H + j0 + u] = __float2bfloat16(h_st[k]);      // h_{-1} -> slot 0
            }
        }
        lt_fence_proxy_async();
        lt_bar(LT_EPI);
        if (tid == 0) lt_publish(bar);

        for (int s = 0; s < p.T; ++s) {
            float gxv[ITEMS][4];
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const int idx = tid + LT_EPI * k, b = idx / FWD_UPC, u = idx % FWD_UPC;
                if (s < len[k]) {
                    const int t = dir == 0 ? s : (len[k] - 1 - s);
                    const float* gxr = p.gx + ((size_t)b * p.T + t) * G + (size_t)dir * 4 * H + j0 + u;
#pragma unroll
                    for (int g = 0; g < 4; ++g) gxv[k][g] = __ldg(gxr + (size_t)g * H);
                }
            }
            tc::mbar_wait(sm.tfull, (uint32_t)(s & 1));
            tc::tc_fence_after();
            lt_tmem_to_smem<M, N>(tmem, sm.pre, B, warp, lane);
            lt_bar(LT_EPI);
            __nv_bfloat16* hnext = hb + (size_t)((s + 1) & 1) * B * H;
            float sv[ITEMS][6];                                   // i, f, g, o, c_{t-1}, h_{t-1} of the items that stepped
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const int idx = tid + LT_EPI * k, b = idx / FWD_UPC, u = idx % FWD_UPC;
                if (b < B) {
                    if (s < len[k]) {
                        const float* pr = sm.pre + b * (N + 1) + u;
                        const float ig = lt_sigmoid(pr[0] + gxv[k][0]);
                        const float fg = lt_sigmoid(pr[FWD_UPC] + gxv[k][1]);
                        const float gg = lt_tanh(pr[2 * FWD_UPC] + gxv[k][2]);
                        const float og = lt_sigmoid(pr[3 * FWD_UPC] + gxv[k][3]);
                        sv[k][0] = ig; sv[k][1] = fg; sv[k][2] = gg; sv[k][3] = og; sv[k][4] = c_st[k]; sv[k][5] = h_st[k];
                        const float c = fg * c_st[k] + ig * gg;
                        c_st[k] = c; h_st[k] = og * lt_tanh(c);
                    }
                    hnext[(size_t)b * H + j0 + u] = __float2bfloat16(h_st[k]);     // a frozen state keeps being republished
                }
            }
            if (!(p.dbg & 4)) lt_fence_proxy_async();
            lt_bar(LT_EPI);
            if (tid == 0) lt_publish(bar);
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const int idx = tid + LT_EPI * k, b = idx / FWD_UPC, u = idx % FWD_UPC;
                if (b < B && s < len[k]) {
                    const int t = dir == 0 ? s : (len[k] - 1 - s);
                    if (p.acts) {
                        const size_t cell = ((size_t)b * p.T + t) * p.ndir + dir;
                        float* ar = p.acts + cell * 4 * H + j0 + u;
                        ar[0] = sv[k][0]; ar[H] = sv[k][1]; ar[2 * (size_t)H] = sv[k][2]; ar[3 * (size_t)H] = sv[k][3];
                        p.cprev[cell * H + j0 + u] = sv[k][4];
                        p.hprev[cell * H + j0 + u] = sv[k][5];
                    }
                    p.y[((size_t)b * p.T + t) * (p.ndir * H) + (size_t)dir * H + j0 + u] = h_st[k];
                }
            }
        }
        if (p.hN) {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const int idx = tid + LT_EPI * k, b = idx / FWD_UPC, u = idx % FWD_UPC;
                if (b < B) {
                    const size_t si = ((size_t)dir * B + b) * H + j0 + u;
                    p.hN[si] = h_st[k];
                    p.cN[si] = c_st[k];
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 5) { tc::tc_fence_after(); tc::tmem_dealloc<LT_TMEM_COLS>(tmem); }
}

template <int M, int BWD_UPC>
__global__ void __launch_bounds__(LT_THREADS, 1) lstm_tc_bwd_kernel(const __grid_constant__ LtBwd p) {
    pdl_entry();
    constexpr int N = BWD_UPC;
    constexpr int ITEMS = M * BWD_UPC / LT_EPI;
    extern __shared__ __align__(1024) uint8_t lt_smem_raw[];
    const int H = p.H, B = p.B, H4 = 4 * p.H;
    const int chunkA = ((B + 7) & ~7) * 128;
    const int kchunks = H4 / 64;
    const LtSmem sm = lt_carve(lt_smem_raw, p.nst, p.ch * chunkA, N * H4 * 2, B * (N + 1));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int per_dir = H / BWD_UPC;
    const int dir = p.dir0 + blockIdx.x / per_dir;
    const int j0 = (blockIdx.x % per_dir) * BWD_UPC;
    const int G = p.ndir * H4;
    unsigned* bar = p.bar + dir;
    const int nmm = p.dh0 ? p.T : p.T - 1;           // nothing flows into a zero initial state

    if (warp == 4 && lane == 0) tc::tma_prefetch_desc(&p.amap);
    if (warp == 5 && lane == 0) {
        for (int i = 0; i < LT_MAXST; ++i) { tc::mbar_init(&sm.full[i], 1); tc::mbar_init(&sm.empty[i], 1); }
        tc::mbar_init(sm.tfull, 1);
        tc::fence_barrier_init();
    }
    if (warp == 5) tc::tmem_alloc<LT_TMEM_COLS>(sm.holder);
    {   // resident B operand: row n = unit, column k = gate row r  <-  W_hh[dir][r][j0 + n], bf16
        const float* wg = p.whh + (size_t)dir * H4 * H;
        const int k8n = H4 / 8;
        for (int e = tid; e < N * k8n; e += LT_THREADS) {
            const int n = e % N, k = (e / N) * 8;
            const float* src = wg + (size_t)k * H + j0 + n;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ldg(src + (size_t)i * H);
            tc::sts128(tc::smem_u32(sm.w) + lt_sw_off(n, k, N * 128), tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]),
                       tc::pack_bf16x2(v[4], v[5]), tc::pack_bf16x2(v[6], v[7]));
        }
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *sm.holder;

    if (warp == 4) {
        if (tc::elect_one())
            lt_producer<M, N>(&p.amap, sm, bar, (unsigned)per_dir, nmm, B, dir * 2, kchunks, p.ch, p.nst, chunkA, p.dbg);
    } else if (warp == 5) {
        if (tc::elect_one()) lt_mma<M, N>(sm, tmem, nmm, kchunks, p.ch, p.nst, chunkA, p.dbg);
    } else {
        float dc_st[ITEMS], dh_rec[ITEMS];
        int len[ITEMS];
        __nv_bfloat16* gb = p.abuf + (size_t)dir * 2 * B * H4;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int b = (tid + LT_EPI * k) / BWD_UPC;
            dc_st[k] = 0.f; dh_rec[k] = 0.f;
            len[k] = b < B ? min(max(p.lens[b], 0), p.T) : 0;
        }
        for (int it = 0; it < p.T; ++it) {
            const int s = p.T - 1 - it;                // forward step index being undone
            float in[ITEMS][6];                        // i, f, g, o, c_{t-1}, dy
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const int idx = tid + LT_EPI * k, b = idx / BWD_UPC, u = idx % BWD_UPC;
                if (s < len[k]) {
                    const int t = dir == 0 ? s : (len[k] - 1 - s);
                    const size_t cell = ((size_t)b * p.T + t) * p.ndir + dir;
                    const float* ar = p.acts + cell * H4 + j0 + u;
#pragma unroll
                    for (int g = 0; g < 4; ++g) in[k][g] = __ldg(ar + (size_t)g * H);
                    in[k][4] = __ldg(p.cprev + cell * H + j0 + u);
                    in[k][5] = __ldg(p.dy + ((size_t)b * p.T + t) * (p.ndir * H) + (size_t)dir * H + j0 + u);
                }
            }
            if (it > 0) {                              // dh_rec of the step undone before this one
                tc::mbar_wait(sm.tfull, (uint32_t)((it - 1) & 1));
                tc::tc_fence_after();
                lt_tmem_to_smem<M, N>(tmem, sm.pre, B, warp, lane);
                lt_bar(LT_EPI);
#pragma unroll
                for (int k = 0; k < ITEMS; ++k) {
                    const int idx = tid + LT_EPI * k, b = idx / BWD_UPC, u = idx % BWD_UPC;
                    if (b < B) dh_rec[k] = sm.pre[b * (N + 1) + u];
                }
            }
            __nv_bfloat16* gslot = gb + (size_t)(it & 1) * B * H4;
            float d4[ITEMS][4];
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const int idx = tid + LT_EPI * k, b = idx / BWD_UPC, u = idx % BWD_UPC;
#pragma unroll
                for (int g = 0; g < 4; ++g) d4[k][g] = 0.f;
                if (b < B) {
                    if (s < len[k]) {
                        const float ig = in[k][0], fg = in[k][1], gg = in[k][2], og = in[k][3], cp = in[k][4];
                        const float c = fg * cp + ig * gg;
                        const float tch = lt_tanh(c);
                        float dh = in[k][5] + dh_rec[k];
                        float dc_in = dc_st[k];
                        if (p.dhN && s == len[k] - 1) {           // the final state is the state after the last valid step
                            const size_t si = ((size_t)dir * B + b) * H + j0 + u;
                            dh += __ldg(p.dhN + si);
                            dc_in += __ldg(p.dcN + si);
                        }
                        const float dc = dc_in + dh * og * (1.f - tch * tch);
                        dc_st[k] = dc * fg;
                        d4[k][0] = dc * gg * ig * (1.f - ig);
                        d4[k][1] = dc * cp * fg * (1.f - fg);
                        d4[k][2] = dc * ig * (1.f - gg * gg);
                        d4[k][3] = dh * tch * og * (1.f - og);
                    }
                    if (it < nmm) {                               // the exchange copy: finished rows publish zeros
                        __nv_bfloat16* gq = gslot + (size_t)b * H4 + j0 + u;
#pragma unroll
                        for (int g = 0; g < 4; ++g) gq[(size_t)g * H] = __float2bfloat16(d4[k][g]);
                    }
                }
            }
            if (it < nmm) {
                if (!(p.dbg & 4)) lt_fence_proxy_async();
                lt_bar(LT_EPI);
                if (tid == 0) lt_publish(bar);
            }
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const int idx = tid + LT_EPI * k, b = idx / BWD_UPC, u = idx % BWD_UPC;
                if (b < B && s < len[k]) {
                    const int t = dir == 0 ? s : (len[k] - 1 - s);
                    float* gr = p.dg + ((size_t)b * p.T + t) * G + (size_t)dir * H4 + j0 + u;
#pragma unroll
                    for (int g = 0; g < 4; ++g) gr[(size_t)g * H] = d4[k][g];
                }
            }
        }
        if (p.dh0) {                                   // gradient w.r.t. (h_0, c_0): what the recurrence passes below step 0
            tc::mbar_wait(sm.tfull, (uint32_t)((p.T - 1) & 1));
            tc::tc_fence_after();
            lt_tmem_to_smem<M, N>(tmem, sm.pre, B, warp, lane);
            lt_bar(LT_EPI);
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const int idx = tid + LT_EPI * k, b = idx / BWD_UPC, u = idx % BWD_UPC;
                if (b < B) {
                    const size_t si = ((size_t)dir * B + b) * H + j0 + u;
                    p.dh0[si] = sm.pre[b * (N + 1) + u];
                    p.dc0[si] = dc_st[k];
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 5) { tc::tc_fence_after(); tc::tmem_dealloc<LT_TMEM_COLS>(tmem); }
}

// shared-memory plan: ring stages that fit next to the resident operand; 0 stages = unsupported
struct LtPlan { int ch, nst; size_t smem; };
LtPlan lt_plan(int B, int K, int N) {
    LtPlan pl{0, 0, 0};
    const int chunkA = ((B + 7) & ~7) * 128;
    const int kchunks = K / 64;
    const size_t fixed = (size_t)N * K * 2 + LT_PAD + (((size_t)B * (N + 1) * 4 + 15) & ~(size_t)15) + (2 * LT_MAXST + 1) * 8 + 16 + 1024;
    const size_t budget = 227 * 1024;
    for (int ch = 4; ch >= 1; ch >>= 1) {
        if (kchunks % ch) continue;
        const size_t stage = (size_t)ch * chunkA;
        if (fixed + 2 * stage > budget) continue;
        int nst = (int)((budget - fixed) / stage);
        nst = nst > LT_MAXST ? LT_MAXST : nst;
        const int nit = kchunks / ch;
        if (nst > nit && nit >= 2) nst = nit;
        if (nst < 2) nst = 2;
        pl.ch = ch; pl.nst = nst; pl.smem = fixed + (size_t)nst * stage;
        return pl;
    }
    return pl;
}

int lt_env(const char* name, int dflt) {          // experiment knobs (profiles/prof_lstm.py); unset in production
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

bool lt_shape_ok(int B, int H, int ndir) {
    if (B <= 0 || B > 128 || H <= 0 || H % 64 != 0 || (ndir != 1 && ndir != 2)) return false;
    if (lt_plan(B, H, 4 * FWD_UPC).nst == 0 || lt_plan(B, 4 * H, BWD_UPC_DEFAULT).nst == 0) return false;
    return H / BWD_UPC_DEFAULT >= 1 && H / FWD_UPC <= num_sms();
}

}  // namespace
}  // namespace nsp

using namespace nsp;

extern "C" int nsp_lstm_tc_supported(int B, int H, int ndir) { return lt_shape_ok(B, H, ndir) ? 1 : 0; }

extern "C" size_t nsp_lstm_tc_workspace_bytes(int B, int H, int ndir, int backward) {
    if (B <= 0 || H <= 0 || ndir <= 0) return 0;
    const size_t K = backward ? (size_t)4 * H : (size_t)H;
    return align_up((size_t)ndir * 2 * B * K * sizeof(__nv_bfloat16), 256) + 256;
}

extern "C" nsp_status nsp_lstm_seq_fwd_tc(const float* gates_x, const float* w_hh, const int32_t* lens, float* y,
                                          int B, int T, int H, int ndir, float* acts, float* cprev, float* hprev,
                                          const float* h0, const float* c0, float* hN, float* cN,
                                          void* workspace, size_t workspace_bytes, void* stream) {
    NSP_CHECK_ARG(gates_x && w_hh && lens && y && workspace, "lstm_seq_fwd_tc: null pointer");
    NSP_CHECK_ARG(B > 0 && T > 0 && H > 0 && (ndir == 1 || ndir == 2), "lstm_seq_fwd_tc: bad shape");
    NSP_CHECK_ARG((acts == nullptr) == (cprev == nullptr) && (acts == nullptr) == (hprev == nullptr), "lstm_seq_fwd_tc: save buffers go together");
    NSP_CHECK_ARG((hN == nullptr) == (cN == nullptr), "lstm_seq_fwd_tc: hN and cN go together");
    if (!lt_shape_ok(B, H, ndir)) { set_error("lstm_seq_fwd_tc: B=%d H=%d unsupported (B <= 128, H %% 64 == 0)", B, H); return NSP_ERR_UNSUPPORTED; }
    NSP_CHECK_ARG(workspace_bytes >= nsp_lstm_tc_workspace_bytes(B, H, ndir, 0), "lstm_seq_fwd_tc: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const LtPlan pl = lt_plan(B, H, 4 * FWD_UPC);
    LtFwd p;
    p.gx = gates_x; p.whh = w_hh; p.lens = lens; p.y = y; p.B = B; p.T = T; p.H = H; p.ndir = ndir;
    p.acts = acts; p.cprev = cprev; p.hprev = hprev; p.h0 = h0; p.c0 = c0; p.hN = hN; p.cN = cN;
    p.ch = pl.ch; p.nst = pl.nst; p.dbg = lt_env("NSP_LSTM_TC_DEBUG", 0);
    const size_t abytes = align_up((size_t)ndir * 2 * B * H * sizeof(__nv_bfloat16), 256);
    p.abuf = (__nv_bfloat16*)workspace;
    p.bar = (unsigned int*)((char*)workspace + abytes);
    if (!encode_tmap_2d(&p.amap, p.abuf, true, (uint64_t)ndir * 2 * B, (uint64_t)H, (uint64_t)H, (uint32_t)B, 64,
                        CU_TENSOR_MAP_SWIZZLE_128B, "lstm h exchange")) return NSP_ERR_INVALID;
    NSP_CUDA_OK(cudaMemsetAsync(p.bar, 0, 256, st));
    NSP_CUDA_OK(cudaMemsetAsync(y, 0, (size_t)B * T * ndir * H * sizeof(float), st));
    const bool m64 = B <= 64 && lt_env("NSP_LSTM_TC_M", 64) != 128;
    void* kern = m64 ? (void*)lstm_tc_fwd_kernel<64> : (void*)lstm_tc_fwd_kernel<128>;
    NSP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    const int per_dir = H / FWD_UPC;
    const int dirs_per_launch = (ndir * per_dir <= num_sms()) ? ndir : 1;
    for (int d0 = 0; d0 < ndir; d0 += dirs_per_launch) {
        p.dir0 = d0;
        void* args[] = {&p};
        NSP_CUDA_OK(cudaLaunchCooperativeKernel(kern, dim3(dirs_per_launch * per_dir), dim3(LT_THREADS), args, pl.smem, st));
    }
    return NSP_OK;
}

extern "C" nsp_status nsp_lstm_seq_bwd_tc(const float* dy, const float* acts, const float* cprev, const float* w_hh,
                                          const int32_t* lens, float* dgates, int B, int T, int H, int ndir,
                                          const float* dhN, const float* dcN, float* dh0, float* dc0,
                                          void* workspace, size_t workspace_bytes, void* stream) {
    NSP_CHECK_ARG(dy && acts && cprev && w_hh && lens && dgates && workspace, "lstm_seq_bwd_tc: null pointer");
    NSP_CHECK_ARG(B > 0 && T > 0 && H > 0 && (ndir == 1 || ndir == 2), "lstm_seq_bwd_tc: bad shape");
    NSP_CHECK_ARG((dhN == nullptr) == (dcN == nullptr) && (dh0 == nullptr) == (dc0 == nullptr), "lstm_seq_bwd_tc: state gradients come in pairs");
    if (!lt_shape_ok(B, H, ndir)) { set_error("lstm_seq_bwd_tc: B=%d H=%d unsupported (B <= 128, H %% 64 == 0)", B, H); return NSP_ERR_UNSUPPORTED; }
    NSP_CHECK_ARG(workspace_bytes >= nsp_lstm_tc_workspace_bytes(B, H, ndir, 1), "lstm_seq_bwd_tc: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    // 8 units per CTA when a direction's H / 8 CTAs fit the GPU (half the resident operand -> twice the ring: the step is bound
    // by the dG stream, 10.1 -> 9.5 us at H = 1024), else 16
    int upc = lt_env("NSP_LSTM_TC_BWD_UPC", H / 8 <= num_sms() ? 8 : 16);
    if (upc != 8 || H / 8 > num_sms() || lt_plan(B, 4 * H, 8).nst == 0) upc = 16;
    LtPlan pl = lt_plan(B, 4 * H, upc);
    if (pl.nst == 0) { set_error("lstm_seq_bwd_tc: no shared-memory plan for B=%d H=%d", B, H); return NSP_ERR_UNSUPPORTED; }
    const int nst_env = lt_env("NSP_LSTM_TC_NST", 0);
    if (nst_env >= 2 && nst_env < pl.nst) { pl.smem -= (size_t)(pl.nst - nst_env) * pl.ch * (((B + 7) & ~7) * 128); pl.nst = nst_env; }
    LtBwd p;
    p.dy = dy; p.acts = acts; p.cprev = cprev; p.whh = w_hh; p.lens = lens; p.dg = dgates;
    p.B = B; p.T = T; p.H = H; p.ndir = ndir; p.dhN = dhN; p.dcN = dcN; p.dh0 = dh0; p.dc0 = dc0;
    p.ch = pl.ch; p.nst = pl.nst; p.dbg = lt_env("NSP_LSTM_TC_DEBUG", 0);
    const size_t abytes = align_up((size_t)ndir * 2 * B * 4 * H * sizeof(__nv_bfloat16), 256);
    p.abuf = (__nv_bfloat16*)workspace;
    p.bar = (unsigned int*)((char*)workspace + abytes);
    if (!encode_tmap_2d(&p.amap, p.abuf, true, (uint64_t)ndir * 2 * B, (uint64_t)4 * H, (uint64_t)4 * H, (uint32_t)B, 64,
                        CU_TENSOR_MAP_SWIZZLE_128B, "lstm dG exchange")) return NSP_ERR_INVALID;
    NSP_CUDA_OK(cudaMemsetAsync(p.bar, 0, 256, st));
    NSP_CUDA_OK(cudaMemsetAsync(dgates, 0, (size_t)B * T * ndir * 4 * H * sizeof(float), st));
    const bool m64 = B <= 64 && lt_env("NSP_LSTM_TC_M", 64) != 128;
    void* kern = upc == 8 ? (m64 ? (void*)lstm_tc_bwd_kernel<64, 8> : (void*)lstm_tc_bwd_kernel<128, 8>)
                          : (m64 ? (void*)lstm_tc_bwd_kernel<64, 16> : (void*)lstm_tc_bwd_kernel<128, 16>);
    NSP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    const int per_dir = H / upc;
    const int dirs_per_launch = (ndir * per_dir <= num_sms()) ? ndir : 1;
    for (int d0 = 0; d0 < ndir; d0 += dirs_per_launch) {
        p.dir0 = d0;
        void* args[] = {&p};
        NSP_CUDA_OK(cudaLaunchCooperativeKernel(kern, dim3(dirs_per_launch * per_dir), dim3(LT_THREADS), args, pl.smem, st));
    }
    return NSP_OK;
}
