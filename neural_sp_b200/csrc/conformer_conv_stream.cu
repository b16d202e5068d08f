// Conformer convolution module core as streaming passes (HBM-bound), forward and backward:
//     y = Swish( LayerNorm( depthwise_conv1d_k(x) + bias ) )          x, y: [B, T, d]  (d contiguous)
// Reference: ConformerConvBlock.forward modules/conformer_convolution.py:113-124 (LayerNorm variant, the LibriSpeech recipes'
// choice); the backward replaces torch autograd over the same lines.
//
// Round 1 staged a (64 + k - 1) x d fp32 window per CTA (190 KB -> one CTA per SM, load phase then compute phase, shared-
// memory atomics for the parameter gradients): 0.15 TB/s on the backward.  Here:
//   conv_ln_kernel<BWD=false>  forward.  A thread owns two adjacent channels and walks 32 + k - 1 frames: all loads of the
//       walk are issued up front (one 32-bit register per bf16 pair), the taps and a k-deep window live in registers, the conv
//       output goes to a 32 x d fp32 tile in shared memory; then each warp normalises 32 / nwarps frames (statistics by
//       shuffles) and writes Swish(LN(z)).  64 KB of shared memory at d = 512 -> several CTAs per SM overlap their phases.
//   conv_ln_kernel<BWD=true>   backward through Swish and LayerNorm: same walk, then dz and per-CTA partial sums of
//       d(norm weight / bias) (registers -> shared memory -> workspace; no atomics).
//   dwconv_bwd_kernel          dx = correlation(dz, taps), d(taps), d(bias): 128-channel x 64-frame tiles of x and dz arrive by
//       cp.async (double buffered); half of the CTA slides a dz window for dx, the other half an x window for d(taps), which
//       stays in registers across the CTA's tiles and leaves as one partial per CTA.
//   conv_bwd_reduce_kernel     sums the partials into dw / dbias / dnorm_w / dnorm_b (+=).
// Shapes outside d in {64, 128, 256, 512}, k in {3, 5, 7, 15}, 16-byte aligned rows fall back to conformer_conv*.cu.
#include "common.cuh"
#include "conv_stream.h"

namespace nsp {
namespace {

constexpr int TT = 32;     // frames per CTA (forward / K1)

template <typename T> struct Raw2;
template <> struct Raw2<__nv_bfloat16> {
    using type = uint32_t;
    static __device__ __forceinline__ type ld(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
    static __device__ __forceinline__ type zero() { return 0u; }
    static __device__ __forceinline__ void unpack(type r, float& a, float& b) { a = __uint_as_float(r << 16); b = __uint_as_float(r & 0xffff0000u); }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, float a, float b) {
        *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
    }
};
template <> struct Raw2<float> {
    using type = float2;
    static __device__ __forceinline__ type ld(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
    static __device__ __forceinline__ type zero() { return make_float2(0.f, 0.f); }
    static __device__ __forceinline__ void unpack(type r, float& a, float& b) { a = r.x; b = r.y; }
    static __device__ __forceinline__ void st(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
};

struct ConvLnParams {
    const void* x; int64_t ldx;
    const float* w;                  // [k][d] taps (transposed)
    const float* bias;               // [d]
    const float* g; const float* b;  // LayerNorm weight / bias
    void* y; int64_t ldy;            // forward: output; backward: dz
    const void* dy; int64_t lddy;    // backward only
    float* part;                     // backward only: [gridDim.x][2][d] partial d(norm weight), d(norm bias)
    int B, T, left_pad;
    float eps;
};

template <typename T, int D, int K, bool BWD>
__global__ void __launch_bounds__(D / 2) conv_ln_kernel(ConvLnParams p) {
    pdl_entry();
    constexpr int NTHR = D / 2;
    constexpr int NW = NTHR / 32;              // warps = 64-channel groups
    constexpr int FPW = TT / NW;               // frames per warp in the normalisation phase
    constexpr int NF = TT + K - 1;             // frames walked by a thread
    constexpr int NB = (sizeof(T) == 2) ? NF : 16;     // frames per load batch (registers: one per frame for bf16 pairs)
    constexpr int CPL = D / 64;                // channel pairs per lane in the normalisation phase
    using R = Raw2<T>;
    extern __shared__ __align__(16) float zs[];        // [TT][D]
    const int ttiles = (p.T + TT - 1) / TT;
    const int b = blockIdx.x / ttiles, t0 = (blockIdx.x % ttiles) * TT;
    const T* xg = reinterpret_cast<const T*>(p.x) + (int64_t)b * p.T * p.ldx;

    // ---------------- phase 1: depthwise convolution, two channels per thread, register window ----------------
    {
        const int c = 2 * threadIdx.x;
        float w0[K], w1[K];
#pragma unroll
        for (int j = 0; j < K; ++j) { const float2 wv = __ldg(reinterpret_cast<const float2*>(p.w + (size_t)j * D + c)); w0[j] = wv.x; w1[j] = wv.y; }
        const float2 bs = __ldg(reinterpret_cast<const float2*>(p.bias + c));
        float x0[K], x1[K];
#pragma unroll
        for (int j = 0; j < K; ++j) { x0[j] = 0.f; x1[j] = 0.f; }
#pragma unroll
        for (int i0 = 0; i0 < NF; i0 += NB) {
            typename R::type raw[NB];
#pragma unroll
            for (int ii = 0; ii < NB; ++ii) {
                const int t = t0 - p.left_pad + i0 + ii;
                raw[ii] = (i0 + ii < NF && t >= 0 && t < p.T) ? R::ld(xg + (int64_t)t * p.ldx + c) : R::zero();
            }
#pragma unroll
            for (int ii = 0; ii < NB; ++ii) {
                if (i0 + ii < NF) {
#pragma unroll
                    for (int j = 0; j < K - 1; ++j) { x0[j] = x0[j + 1]; x1[j] = x1[j + 1]; }
                    R::unpack(raw[ii], x0[K - 1], x1[K - 1]);
                    const int r = i0 + ii - (K - 1);
                    if (r >= 0) {
                        float a0 = bs.x, a1 = bs.y;
#pragma unroll
                        for (int j = 0; j < K; ++j) { a0 = fmaf(w0[j], x0[j], a0); a1 = fmaf(w1[j], x1[j], a1); }
                        *reinterpret_cast<float2*>(zs + (size_t)r * D + c) = make_float2(a0, a1);
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---------------- phase 2: per-frame LayerNorm (+ Swish forward | Swish' and LayerNorm backward) ----------------
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float gm[CPL][2], bt[CPL][2];
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        const int c = 2 * lane + 64 * i;
        const float2 gv = __ldg(reinterpret_cast<const float2*>(p.g + c)), bv = __ldg(reinterpret_cast<const float2*>(p.b + c));
        gm[i][0] = gv.x; gm[i][1] = gv.y; bt[i][0] = bv.x; bt[i][1] = bv.y;
    }
    float dgs[CPL][2], dbs[CPL][2];
    if constexpr (BWD) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) { dgs[i][0] = dgs[i][1] = dbs[i][0] = dbs[i][1] = 0.f; }
    }
    T* yg = reinterpret_cast<T*>(p.y) + (int64_t)b * p.T * p.ldy;
    const T* dyg = BWD ? reinterpret_cast<const T*>(p.dy) + (int64_t)b * p.T * p.lddy : nullptr;
#pragma unroll 1
    for (int f = 0; f < FPW; ++f) {
        const int r = warp * FPW + f, t = t0 + r;
        if (t >= p.T) break;                               // warp-uniform
        typename R::type dyr[CPL];
        if constexpr (BWD) {
#pragma unroll
            for (int i = 0; i < CPL; ++i) dyr[i] = R::ld(dyg + (int64_t)t * p.lddy + 2 * lane + 64 * i);
        }
        float z[CPL][2];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const float2 v = *reinterpret_cast<const float2*>(zs + (size_t)r * D + 2 * lane + 64 * i);
            z[i][0] = v.x; z[i][1] = v.y;
            s += v.x + v.y;
        }
        const float mean = warp_sum(s) * (1.f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) { const float d0 = z[i][0] - mean, d1 = z[i][1] - mean; q += d0 * d0 + d1 * d1; }
        const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + p.eps);
        if constexpr (!BWD) {
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                float o[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float v = (z[i][e] - mean) * rstd * gm[i][e] + bt[i][e];
                    o[e] = __fdividef(v, 1.f + __expf(-v));     // Swish
                }
                R::st(yg + (int64_t)t * p.ldy + 2 * lane + 64 * i, o[0], o[1]);
            }
        } else {
            float dn[CPL][2], nn[CPL][2];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                float dyv[2];
                R::unpack(dyr[i], dyv[0], dyv[1]);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float n = (z[i][e] - mean) * rstd;
                    const float a = n * gm[i][e] + bt[i][e];
                    const float sg = 1.f / (1.f + __expf(-a));
                    const float da = dyv[e] * sg * (1.f + a * (1.f - sg));
                    dgs[i][e] += da * n;
                    dbs[i][e] += da;
                    nn[i][e] = n;
                    dn[i][e] = da * gm[i][e];
                    s1 += dn[i][e];
                    s2 += dn[i][e] * n;
                }
            }
            s1 = warp_sum(s1) * (1.f / D);
            s2 = warp_sum(s2) * (1.f / D);
#pragma unroll
            for (int i = 0; i < CPL; ++i)
                R::st(yg + (int64_t)t * p.ldy + 2 * lane + 64 * i, rstd * (dn[i][0] - s1 - nn[i][0] * s2),
                      rstd * (dn[i][1] - s1 - nn[i][1] * s2));
        }
    }
    if constexpr (BWD) {
        // per-CTA partial of d(norm weight) / d(norm bias): warps -> shared memory (the z tile is dead) -> workspace
        __syncthreads();
        float* red = zs;                                   // [NW][2][D]
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int c = 2 * lane + 64 * i;
            *reinterpret_cast<float2*>(red + ((size_t)warp * 2 + 0) * D + c) = make_float2(dgs[i][0], dgs[i][1]);
            *reinterpret_cast<float2*>(red + ((size_t)warp * 2 + 1) * D + c) = make_float2(dbs[i][0], dbs[i][1]);
        }
        __syncthreads();
        float* out = p.part + (size_t)blockIdx.x * 2 * D;
        for (int e = threadIdx.x; e < 2 * D; e += NTHR) {
            float a = 0.f;
#pragma unroll
            for (int w_ = 0; w_ < NW; ++w_) a += red[(size_t)w_ * 2 * D + e];
            out[e] = a;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
constexpr int K2_CH = 128;      // channels per CTA
constexpr int K2_TC = 64;       // frames per tile
constexpr int K2_NT = 256;

struct DwBwdParams {
    const void* x; int64_t ldx;
    const void* dz; int64_t lddz;
    const float* w;                // [k][d]
    void* dx; int64_t lddx;
    float* part;                   // [gridDim.x][k + 1][K2_CH]  d(taps) rows then the d(bias) row, per CTA
    int B, T, d, left_pad;
    int seg_tiles;                 // tiles per CTA along time
    int nseg;                      // segments per utterance
};

__device__ __forceinline__ void cs_cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
template <typename T> __device__ __forceinline__ float ld_s(const T* p);
template <> __device__ __forceinline__ float ld_s<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float ld_s<float>(const float* p) { return *p; }
template <typename T> __device__ __forceinline__ void st_g(T* p, float v);
template <> __device__ __forceinline__ void st_g<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void st_g<float>(float* p, float v) { *p = v; }

template <typename T, int K>
__global__ void __launch_bounds__(K2_NT) dwconv_bwd_kernel(DwBwdParams p) {
    pdl_entry();
    constexpr int ROWS = K2_TC + K - 1;
    constexpr int CHUNKS = K2_CH * (int)sizeof(T) / 16;      // 16-byte chunks per tile row
    extern __shared__ __align__(16) uint8_t k2_smem[];
    // stage s: xs [ROWS][K2_CH], zs [ROWS][K2_CH]
    auto xs_of = [&](int s) { return reinterpret_cast<T*>(k2_smem) + (size_t)s * 2 * ROWS * K2_CH; };
    auto zs_of = [&](int s) { return xs_of(s) + (size_t)ROWS * K2_CH; };
    const int nslices = p.d / K2_CH;
    const int slice = blockIdx.x % nslices;
    const int seg = (blockIdx.x / nslices) % p.nseg;
    const int b = blockIdx.x / (nslices * p.nseg);
    const int c0 = slice * K2_CH;
    const int rpad = K - 1 - p.left_pad;
    const int ttiles = (p.T + K2_TC - 1) / K2_TC;
    const int tile_lo = seg * p.seg_tiles, tile_hi = min(ttiles, tile_lo + p.seg_tiles);
    const T* xg = reinterpret_cast<const T*>(p.x) + (int64_t)b * p.T * p.ldx + c0;
    const T* zg = reinterpret_cast<const T*>(p.dz) + (int64_t)b * p.T * p.lddz + c0;
    T* dxg = reinterpret_cast<T*>(p.dx) + (int64_t)b * p.T * p.lddx + c0;

    auto stage = [&](int tile, int s) {
        const int t0 = tile * K2_TC;
        T* xs = xs_of(s);
        T* zs = zs_of(s);
        for (int e = threadIdx.x; e < 2 * ROWS * CHUNKS; e += K2_NT) {
            const int which = e / (ROWS * CHUNKS), rem = e % (ROWS * CHUNKS);
            const int r = rem / CHUNKS, ch = rem % CHUNKS;
            const int t = which ? (t0 - rpad + r) : (t0 - p.left_pad + r);
            T* dst = (which ? zs : xs) + (size_t)r * K2_CH + ch * (16 / sizeof(T));
            if (t >= 0 && t < p.T) {
                const T* src = (which ? zg + (int64_t)t * p.lddz : xg + (int64_t)t * p.ldx) + ch * (16 / sizeof(T));
                cs_cp_async16(dst, src);
            } else {
                *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int c = threadIdx.x & (K2_CH - 1);
    const bool do_dx = threadIdx.x < K2_CH;                  // warps 0-3: dx ; warps 4-7: d(taps), d(bias)
    float wj[K], acc[K];
    float dbacc = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) { wj[j] = do_dx ? __ldg(p.w + (size_t)j * p.d + c0 + c) : 0.f; acc[j] = 0.f; }

    if (tile_lo < tile_hi) stage(tile_lo, 0);
    for (int tile = tile_lo; tile < tile_hi; ++tile) {
        const int s = (tile - tile_lo) & 1;
        if (tile + 1 < tile_hi) { stage(tile + 1, s ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const T* xs = xs_of(s);
        const T* zs = zs_of(s);
        const int t0 = tile * K2_TC;
        const int nfr = min(K2_TC, p.T - t0);
        if (do_dx) {
            // dx[t0 + r] = sum_j w[j] * dz[t0 + r + left_pad - j]  ;  dz tile row i <-> frame t0 - rpad + i  ->  i = r + K - 1 - j
            float win[K];
#pragma unroll
            for (int j = 0; j < K - 1; ++j) win[j + 1] = ld_s<T>(zs + (size_t)j * K2_CH + c);
            for (int r = 0; r < nfr; ++r) {
#pragma unroll
                for (int j = 0; j < K - 1; ++j) win[j] = win[j + 1];
                win[K - 1] = ld_s<T>(zs + (size_t)(r + K - 1) * K2_CH + c);
                float a = 0.f;
#pragma unroll
                for (int j = 0; j < K; ++j) a = fmaf(wj[j], win[K - 1 - j], a);
                st_g<T>(dxg + (int64_t)(t0 + r) * p.lddx + c, a);
            }
        } else {
            // d(taps)[j] += dz[t'] * x[t' - left_pad + j]  ;  x tile row i <-> frame t0 - left_pad + i  ->  i = r + j ;  dz row r + rpad
            float win[K];
#pragma unroll
            for (int j = 0; j < K - 1; ++j) win[j + 1] = ld_s<T>(xs + (size_t)j * K2_CH + c);
            for (int r = 0; r < nfr; ++r) {
#pragma unroll
                for (int j = 0; j < K - 1; ++j) win[j] = win[j + 1];
                win[K - 1] = ld_s<T>(xs + (size_t)(r + K - 1) * K2_CH + c);
                const float dzv = ld_s<T>(zs + (size_t)(r + rpad) * K2_CH + c);
                dbacc += dzv;
#pragma unroll
                for (int j = 0; j < K; ++j) acc[j] = fmaf(dzv, win[j], acc[j]);
            }
        }
        __syncthreads();                                    // stage s is refilled by the next iteration's prefetch
    }
    if (!do_dx) {
        float* out = p.part + (size_t)blockIdx.x * (K + 1) * K2_CH;
#pragma unroll
        for (int j = 0; j < K; ++j) out[(size_t)j * K2_CH + c] = acc[j];
        out[(size_t)K * K2_CH + c] = dbacc;
    }
}

// partial sums -> parameter gradients (+=).  part1: [n1][2][d] (d norm weight, d norm bias); part2: [b][seg][slice][k+1][128].
struct ReduceParams {
    const float* part1; int n1;
    const float* part2; int n2;    // CTAs per slice = B * nseg
    float* dw; float* dbias; float* dg; float* db;
    int d, k;
};
__global__ void __launch_bounds__(256) conv_bwd_reduce_kernel(ReduceParams p) {
    pdl_entry();
    const int e = blockIdx.x * 256 + threadIdx.x;
    const int n_ln = 2 * p.d, n_dw = (p.k + 1) * p.d;
    if (e < n_ln) {
        if (p.part1 == nullptr) return;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int i = 0;
        for (; i + 4 <= p.n1; i += 4) {
            a0 += p.part1[(size_t)(i + 0) * n_ln + e]; a1 += p.part1[(size_t)(i + 1) * n_ln + e];
            a2 += p.part1[(size_t)(i + 2) * n_ln + e]; a3 += p.part1[(size_t)(i + 3) * n_ln + e];
        }
        for (; i < p.n1; ++i) a0 += p.part1[(size_t)i * n_ln + e];
        const float a = (a0 + a1) + (a2 + a3);
        if (e < p.d) { if (p.dg) p.dg[e] += a; }
        else if (p.db) p.db[e - p.d] += a;
    } else if (e < n_ln + n_dw) {
        const int q = e - n_ln;
        const int j = q / p.d, ch = q % p.d;               // j == k: the bias row
        const int nslices = p.d / K2_CH;
        const int slice = ch / K2_CH, cc = ch % K2_CH;
        float a0 = 0.f, a1 = 0.f;
        int i = 0;
        for (; i + 2 <= p.n2; i += 2) {
            a0 += p.part2[(((size_t)(i + 0) * nslices + slice) * (p.k + 1) + j) * K2_CH + cc];
            a1 += p.part2[(((size_t)(i + 1) * nslices + slice) * (p.k + 1) + j) * K2_CH + cc];
        }
        for (; i < p.n2; ++i) a0 += p.part2[(((size_t)i * nslices + slice) * (p.k + 1) + j) * K2_CH + cc];
        const float a = a0 + a1;
        if (j < p.k) { if (p.dw) p.dw[(size_t)j * p.d + ch] += a; }
        else if (p.dbias) p.dbias[ch] += a;
    }
}

template <typename T, int D, int K, bool BWD>
nsp_status launch_ln(const ConvLnParams& p, cudaStream_t st) {
    auto kern = conv_ln_kernel<T, D, K, BWD>;
    constexpr size_t smem = sizeof(float) * (size_t)TT * D;
    if (smem > 48 * 1024) {
        static bool done = false;
        if (!done) { NSP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); done = true; }
    }
    launch_k(kern, dim3((unsigned)(p.B * ceil_div(p.T, TT))), dim3(D / 2), smem, st, p);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

template <typename T, bool BWD>
nsp_status dispatch_ln(const ConvLnParams& p, int d, int k, cudaStream_t st) {
#define NSP_LN(DD, KK) if (d == DD && k == KK) return launch_ln<T, DD, KK, BWD>(p, st)
    NSP_LN(512, 15); NSP_LN(512, 7); NSP_LN(512, 5); NSP_LN(512, 3);
    NSP_LN(256, 15); NSP_LN(256, 7); NSP_LN(256, 5); NSP_LN(256, 3);
    NSP_LN(128, 15); NSP_LN(128, 7); NSP_LN(128, 5); NSP_LN(128, 3);
    NSP_LN(64, 15); NSP_LN(64, 7); NSP_LN(64, 5); NSP_LN(64, 3);
#undef NSP_LN
    return NSP_ERR_UNSUPPORTED;
}

template <typename T, int K>
nsp_status launch_dw(const DwBwdParams& p, unsigned grid, cudaStream_t st) {
    auto kern = dwconv_bwd_kernel<T, K>;
    constexpr size_t smem = (size_t)2 * 2 * (K2_TC + K - 1) * K2_CH * sizeof(T);
    if (smem > 48 * 1024) {
        static bool done = false;
        if (!done) { NSP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); done = true; }
    }
    launch_k(kern, dim3(grid), dim3(K2_NT), smem, st, p);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

bool shape_ok(int d, int k) {
    return (d == 64 || d == 128 || d == 256 || d == 512) && (k == 3 || k == 5 || k == 7 || k == 15);
}
bool aligned(const void* ptr, int64_t ld, size_t esz) { return ((uintptr_t)ptr % 16 == 0) && ((ld * (int64_t)esz) % 16 == 0); }

void dw_geometry(int B, int T, int d, int* seg_tiles, int* nseg) {
    // CTAs = B * nseg * (d / 128); aim for >= 2 x SMs while keeping segments as long as possible
    const int ttiles = ceil_div(T, K2_TC);
    const int nslices = d / K2_CH > 0 ? d / K2_CH : 1;
    int want = ceil_div(2 * num_sms(), B * nslices);
    if (want < 1) want = 1;
    if (want > ttiles) want = ttiles;
    *seg_tiles = ceil_div(ttiles, want);
    *nseg = ceil_div(ttiles, *seg_tiles);
}

}  // namespace

size_t conv_stream_bwd_workspace_bytes(int B, int T, int d, int k) {
    if (B <= 0 || T <= 0 || d <= 0 || k <= 0) return 0;
    int seg_tiles, nseg;
    dw_geometry(B, T, d, &seg_tiles, &nseg);
    const size_t n1 = (size_t)B * ceil_div(T, TT) * 2 * d;
    const size_t n2 = (size_t)B * nseg * (d / K2_CH > 0 ? d / K2_CH : 1) * (k + 1) * K2_CH;
    return (n1 + n2) * sizeof(float) + 512;
}

nsp_status conv_stream_fwd(int is_bf16, const void* x, int64_t ldx, const float* w, const float* bias, const float* g,
                           const float* bta, float eps, void* y, int64_t ldy, int B, int T, int d, int k, int causal,
                           cudaStream_t st) {
    const size_t esz = is_bf16 ? 2 : 4;
    if (!shape_ok(d, k) || !aligned(x, ldx, esz) || !aligned(y, ldy, esz) || ((uintptr_t)w % 8) || ((uintptr_t)bias % 8) ||
        ((uintptr_t)g % 8) || ((uintptr_t)bta % 8))
        return NSP_ERR_UNSUPPORTED;
    ConvLnParams p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.ldx = ldx; p.w = w; p.bias = bias; p.g = g; p.b = bta; p.y = y; p.ldy = ldy;
    p.B = B; p.T = T; p.left_pad = causal ? (k - 1) : (k - 1) / 2; p.eps = eps;
    return is_bf16 ? dispatch_ln<__nv_bfloat16, false>(p, d, k, st) : dispatch_ln<float, false>(p, d, k, st);
}

nsp_status conv_stream_bwd(int is_bf16, const void* x, int64_t ldx, const float* w, const float* bias, const float* g,
                           const float* bta, float eps, const void* dy, int64_t lddy, void* dz, int64_t lddz, void* dx,
                           int64_t lddx, float* dw, float* dbias, float* dg, float* db, int B, int T, int d, int k, int causal,
                           void* ws, size_t ws_bytes, cudaStream_t st) {
    const size_t esz = is_bf16 ? 2 : 4;
    if (!shape_ok(d, k) || d < K2_CH || !aligned(x, ldx, esz) || !aligned(dy, lddy, esz) || !aligned(dz, lddz, esz) ||
        !aligned(dx, lddx, esz) || ((uintptr_t)w % 8) || ((uintptr_t)bias % 8) || ((uintptr_t)g % 8) || ((uintptr_t)bta % 8) ||
        ws == nullptr || ws_bytes < conv_stream_bwd_workspace_bytes(B, T, d, k))
        return NSP_ERR_UNSUPPORTED;
    float* part1 = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    const int n1 = B * ceil_div(T, TT);
    float* part2 = part1 + (size_t)n1 * 2 * d;
    ConvLnParams p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.ldx = ldx; p.w = w; p.bias = bias; p.g = g; p.b = bta; p.y = dz; p.ldy = lddz; p.dy = dy; p.lddy = lddy;
    p.part = part1; p.B = B; p.T = T; p.left_pad = causal ? (k - 1) : (k - 1) / 2; p.eps = eps;
    nsp_status s = is_bf16 ? dispatch_ln<__nv_bfloat16, true>(p, d, k, st) : dispatch_ln<float, true>(p, d, k, st);
    if (s != NSP_OK) return s;
    DwBwdParams q;
    memset(&q, 0, sizeof(q));
    q.x = x; q.ldx = ldx; q.dz = dz; q.lddz = lddz; q.w = w; q.dx = dx; q.lddx = lddx; q.part = part2;
    q.B = B; q.T = T; q.d = d; q.left_pad = p.left_pad;
    dw_geometry(B, T, d, &q.seg_tiles, &q.nseg);
    const unsigned grid = (unsigned)(B * q.nseg * (d / K2_CH));
#define NSP_DW(KK) if (k == KK) s = is_bf16 ? launch_dw<__nv_bfloat16, KK>(q, grid, st) : launch_dw<float, KK>(q, grid, st)
    NSP_DW(15); NSP_DW(7); NSP_DW(5); NSP_DW(3);
#undef NSP_DW
    if (s != NSP_OK) return s;
    ReduceParams r;
    r.part1 = part1; r.n1 = n1; r.part2 = part2; r.n2 = B * q.nseg; r.dw = dw; r.dbias = dbias; r.dg = dg; r.db = db; r.d = d; r.k = k;
    launch_k(conv_bwd_reduce_kernel, dim3((unsigned)ceil_div(2 * d + (k + 1) * d, 256)), dim3(256), 0, st, r);
    NSP_LAUNCH_OK();
    return NSP_OK;
}

}  // namespace nsp
