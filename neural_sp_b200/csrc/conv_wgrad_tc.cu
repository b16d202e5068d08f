// Weight gradient of the 32 -> 32 channel 3x3 convolutions of the front-end as an implicit GEMM on tcgen05.
//
// The reference obtains it from autograd over nn.Conv2d(32, 32, 3, padding 1) in Conv2dBlock.forward (encoders/conv.py:362-394):
//     dW[co, ci, ky, kx] += sum_{b,t,f} dz[b,t,f,co] * a[b, t+ky-1, f+kx-1, ci]
// GEMM view: M = (tap, ci) = 288 rows in three groups of 128 (4 taps x 32 channels; the last group holds one tap),
// N = co = 32, K = positions.  Both operands are read position-major exactly as they sit in HBM: per 8x16-position tile
// nine shifted TMA boxes of the activations (zero-filled outside the tensor = the convolution's padding) and one box of
// dz land as 128-row x 64-byte, 64B-swizzled tiles, which are MN-major tcgen05 operands (LBO = one tap tile, SBO = 8
// positions).  Persistent CTAs accumulate their tiles in TMEM ([128 x 32] x 3) and add the result to dW once.
#include "tc_common.cuh"

namespace nsp {
namespace {

constexpr int TT = 8, TF = 16;              // tile: frames x bins = 128 positions
constexpr int BOX = TT * TF * 64;           // 8 KiB per tap tile
constexpr int STAGE = 10 * BOX;             // 9 activation taps + dz
constexpr int NSTG = 2;

__device__ __forceinline__ uint64_t desc_mn_sw64(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // stride between 32-element (64 B) chunks of the M / N index
    d |= (uint64_t)(512 >> 4) << 32;                     // 8 positions x 64 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                              // SWIZZLE_64B
    return d;
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :: "r"(tc::smem_u32(smem_dst)), "l"(m), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__global__ void __launch_bounds__(192, 1) conv3x3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                   const __grid_constant__ CUtensorMap tmap_z,
                                                                   float* __restrict__ dw, int B, int T, int F) {
    pdl_launch_dependents();      // PDL: the next kernel may start its prologue; ours overlaps the previous kernel's tail
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sS = smem;                                               // NSTG stages, then 2 boxes of slack for group 2
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTG * STAGE + 2 * BOX);
    uint64_t* empty_bar = full_bar + NSTG;
    uint64_t* tfull_bar = empty_bar + NSTG;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tfull_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_f = (F + TF - 1) / TF, tiles_t = (T + TT - 1) / TT;
    const int num_tiles = B * tiles_t * tiles_f;

    if (warp == 0 && lane == 0) { tc::tma_prefetch_desc(&tmap_a); tc::tma_prefetch_desc(&tmap_z); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NSTG; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
        tc::mbar_init(tfull_bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<128>(tmem_holder);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();                    // the previous grid is complete: operands / residuals / outputs may be touched from here

    if (warp == 0) {
        if (tc::elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int ft = tile % tiles_f, tt = (tile / tiles_f) % tiles_t, b = tile / (tiles_f * tiles_t);
                tc::mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* st = sS + stage * STAGE;
                tc::mbar_arrive_expect_tx(&full_bar[stage], STAGE);
                for (int tap = 0; tap < 9; ++tap)
                    tma_load_4d(st + tap * BOX, &tmap_a, &full_bar[stage], 0, ft * TF + (tap % 3) - 1, tt * TT + (tap / 3) - 1, b);
                tma_load_4d(st + 9 * BOX, &tmap_z, &full_bar[stage], 0, ft * TF, tt * TT, b);
                if (++stage == NSTG) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::make_idesc(1u, 128, 32) | (1u << 15) | (1u << 16);     // A and B MN-major
            int stage = 0; uint32_t phase = 0; int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                tc::mbar_wait(&full_bar[stage], phase);
                tc::tc_fence_after();
                const uint32_t st = tc::smem_u32(sS + stage * STAGE);
                const uint64_t bdesc = desc_mn_sw64(st + 9 * BOX, 16);
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    // rows of this group = taps 4g .. 4g+3 (32 channels each), BOX bytes apart; the rows of taps >= 9 read
                    // whatever follows in shared memory and are never stored
                    const uint64_t adesc = desc_mn_sw64(st + g * 4 * BOX, BOX);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {                     // 16 positions = 1024 B per MMA
                        const uint64_t koff = (uint64_t)((ks * 1024) >> 4);
                        tc::umma_f16(tmem_base + (uint32_t)(g * 32), adesc + koff, bdesc + koff, idesc, (it > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                tc::umma_commit(&empty_bar[stage]);
                if (++stage == NSTG) { stage = 0; phase ^= 1; }
            }
            tc::umma_commit(tfull_bar);
        }
    } else {
        const int q = warp & 3;
        if (blockIdx.x < num_tiles) {
            tc::mbar_wait_epi(tfull_bar, 0);
            tc::tc_fence_after();
#pragma unroll 1
            for (int g = 0; g < 3; ++g) {
                const int tap = 4 * g + q;
                uint32_t r[32];
                tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 32), r);
                tc::tmem_ld_wait();
                if (tap < 9) {
#pragma unroll
                    for (int co = 0; co < 32; ++co) atomicAdd(dw + ((int64_t)co * 32 + lane) * 9 + tap, __uint_as_float(r[co]));
                }
            }
            tc::tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc<128>(tmem_base); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

bool get_tma_encode(void** fn);   // gemm_tcgen05.cu

}  // namespace nsp

using namespace nsp;

// a, dz: bf16 [B,T,F,32] channels-last; dw: fp32 [32,32,3,3] (accumulated)
extern "C" nsp_status nsp_conv3x3_c32_wgrad_tc(const void* a, const void* dz, float* dw, int B, int T, int F, void* stream) {
    NSP_CHECK_ARG(a && dz && dw, "conv3x3_wgrad_tc: null pointer");
    NSP_CHECK_ARG(B > 0 && T > 0 && F > 0, "conv3x3_wgrad_tc: bad shape");
    NSP_CHECK_ARG(((uintptr_t)a % 16 == 0) && ((uintptr_t)dz % 16 == 0), "conv3x3_wgrad_tc: alignment");
    void* fnp = nullptr;
    if (!get_tma_encode(&fnp)) return NSP_ERR_CUDA;
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    CUtensorMap ta, tz;
    const void* bases[2] = {a, dz};
    CUtensorMap* maps[2] = {&ta, &tz};
    for (int i = 0; i < 2; ++i) {
        cuuint64_t gdim[4] = {32, (cuuint64_t)F, (cuuint64_t)T, (cuuint64_t)B};
        cuuint64_t gstr[3] = {64, (cuuint64_t)F * 64, (cuuint64_t)T * F * 64};
        cuuint32_t box[4] = {32, TF, TT, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(bases[i]), gdim, gstr, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3x3_wgrad_tc: tensor map failed: %d", (int)r); return NSP_ERR_CUDA; }
    }
    const size_t smem = 1024 + (size_t)NSTG * STAGE + 2 * BOX + 256;
    static bool attr = false;
    if (!attr) { NSP_CUDA_OK(cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    const int tiles = B * ceil_div(T, TT) * ceil_div(F, TF);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    launch_k(conv3x3_wgrad_tc_kernel, dim3(grid), dim3(192), smem, (cudaStream_t)stream, ta, tz, dw, B, T, F);
    NSP_LAUNCH_OK();
    return NSP_OK;
}
