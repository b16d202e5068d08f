// 3x3 convolution (C_in = C_out = 32) + bias + ReLU (+ 2x2 ceil-mode max-pool) as an implicit GEMM on tcgen05.
//
// Replaces  nn.Conv2d(32, 32, 3, padding 1) + ReLU (+ nn.MaxPool2d((2,2), ceil_mode=True)) of
//           Conv2dBlock.forward  encoders/conv.py:362-394  for the 32_32 front-end of the LibriSpeech recipes.
//
// Activations are channels-last bf16 [B, T, F, 32]: one position = one 64-byte row.  A CTA computes a tile of
// 8 frames x 16 bins = 128 positions (the MMA's M) x 32 output channels (N).  K = 9 taps x 32 channels: for each
// tap one TMA 4-D box [1, 8, 16, 32ch] shifted by (ky-1, kx-1) lands as a 128-row x 64-byte, 64B-swizzled A tile
// (out-of-range coordinates are zero-filled by TMA = the convolution's zero padding), and two tcgen05.mma
// (M128 N32 K16) per tap accumulate into TMEM.  The 18 KiB of weights ([32][tap*32 + ci], K-major) stay resident
// in shared memory.  Warp roles and the mbarrier pipelines are those of gemm_tcgen05.cu.
// Roofline: the 9 shifted re-reads of each input tile hit L2, so the kernel is L2->SMEM bound (TMA), not MMA bound.
#include "tc_common.cuh"

namespace nsp {
namespace {

constexpr int TT = 8, TF = 16;              // tile: frames x bins
constexpr int A_BYTES = TT * TF * 64;       // 8 KiB per tap tile
constexpr int W_TAP_BYTES = 32 * 64;        // 2 KiB per tap
constexpr int NST = 16;                     // A ring depth (taps in flight)

struct ConvTcArgs {
    const float* bias;                      // [32]
    __nv_bfloat16* y;                       // [B, To, Fo, 32]
    int B, T, F, To, Fo;
    int pool;                               // 1: fused 2x2 ceil-mode max-pool
    int relu;
};

__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;        // 8 rows x 64 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                 // SWIZZLE_64B
    return d;
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :: "r"(tc::smem_u32(smem_dst)), "l"(m), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__global__ void __launch_bounds__(192, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                             const __grid_constant__ CUtensorMap tmap_w,
                                                             const ConvTcArgs a) {
    pdl_launch_dependents();      // PDL: the next kernel may start its prologue; ours overlaps the previous kernel's tail
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sW = smem;                                   // 9 x 2 KiB
    uint8_t* sA = smem + 9 * W_TAP_BYTES + 2048;          // keep 1024-alignment: 18 KiB + 2 KiB pad = 20 KiB
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sA + NST * A_BYTES);
    uint64_t* empty_bar = full_bar + NST;
    uint64_t* tfull_bar = empty_bar + NST;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* w_bar = tempty_bar + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(w_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_f = (a.F + TF - 1) / TF, tiles_t = (a.T + TT - 1) / TT;
    const int num_tiles = a.B * tiles_t * tiles_f;

    if (warp == 0 && lane == 0) { tc::tma_prefetch_desc(&tmap_x); tc::tma_prefetch_desc(&tmap_w); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NST; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull_bar[i], 1); tc::mbar_init(&tempty_bar[i], 4); }
        tc::mbar_init(w_bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<64>(tmem_holder);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();                    // the previous grid is complete: operands / residuals / outputs may be touched from here

    if (warp == 0) {
        if (tc::elect_one()) {
            tc::mbar_arrive_expect_tx(w_bar, 9 * W_TAP_BYTES);
            for (int tap = 0; tap < 9; ++tap) tc::tma_load_2d(sW + tap * W_TAP_BYTES, &tmap_w, w_bar, tap * 32, 0);
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int ft = tile % tiles_f, tt = (tile / tiles_f) % tiles_t, b = tile / (tiles_f * tiles_t);
                for (int tap = 0; tap < 9; ++tap) {
                    tc::mbar_wait_spin(&empty_bar[stage], phase ^ 1);
                    tc::mbar_arrive_expect_tx(&full_bar[stage], A_BYTES);
                    tma_load_4d(sA + stage * A_BYTES, &tmap_x, &full_bar[stage], 0, ft * TF + (tap % 3) - 1,
                                tt * TT + (tap / 3) - 1, b);
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::make_idesc(1u, 128, 32);
            tc::mbar_wait_spin(w_bar, 0);
            tc::tc_fence_after();
            int stage = 0; uint32_t phase = 0; int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                tc::mbar_wait_spin(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * 32);
                for (int tap = 0; tap < 9; ++tap) {
                    tc::mbar_wait_spin(&full_bar[stage], phase);
                    tc::tc_fence_after();
                    const uint64_t adesc = make_smem_desc_sw64(tc::smem_u32(sA + stage * A_BYTES));
                    const uint64_t bdesc = make_smem_desc_sw64(tc::smem_u32(sW + tap * W_TAP_BYTES));
                    tc::umma_f16(d_tmem, adesc, bdesc, idesc, tap > 0 ? 1u : 0u);
                    tc::umma_f16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);      // +32 bytes of K
                    tc::umma_commit(&empty_bar[stage]);
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
                tc::umma_commit(&tfull_bar[as]);
            }
        }
    } else {
        const int q = warp & 3;
        float bias[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) bias[j] = __ldg(a.bias + j);
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int ft = tile % tiles_f, tt = (tile / tiles_f) % tiles_t, b = tile / (tiles_f * tiles_t);
            const int as = it & 1;
            tc::mbar_wait_epi(&tfull_bar[as], (it >> 1) & 1);
            tc::tc_fence_after();
            uint32_t r[32];
            tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 32), r);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty_bar[as]);          // accumulator drained to registers
            const int row = q * 32 + lane;
            const int t = tt * TT + row / TF, f = ft * TF + row % TF;
            const bool valid = (t < a.T) && (f < a.F);
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(r[j]) + bias[j];
                if (a.relu) x = fmaxf(x, 0.f);
                v[j] = valid ? x : -INFINITY;
            }
            bool writer = valid;
            int to = t, fo = f;
            if (a.pool) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));     // bins f, f^1
                    v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 16));    // frames t, t^1 (same warp: 2 frames x 16 bins)
                }
                writer = valid && ((lane & 1) == 0) && ((lane & 16) == 0);
                to = t >> 1; fo = f >> 1;
            }
            if (writer) {
                __nv_bfloat16* o = a.y + (((int64_t)b * a.To + to) * a.Fo + fo) * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[j], v[j + 1]), p1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), p3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                    uint4 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
                    pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
                    *reinterpret_cast<uint4*>(o + j) = pk;
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc<64>(tmem_base); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

bool get_tma_encode(void** fn);   // gemm_tcgen05.cu

}  // namespace nsp

using namespace nsp;

// x: bf16 [B,T,F,32] channels-last; w_taps: bf16 [32, 288] with column index tap*32 + ci (tap = ky*3 + kx);
// y: bf16 [B,To,Fo,32] with (To,Fo) = (T,F) or (ceil(T/2), ceil(F/2)) when pool.
extern "C" nsp_status nsp_conv3x3_c32_tc_fwd(const void* x, const void* w_taps, const float* bias, void* y,
                                             int B, int T, int F, int relu, int pool2x2, void* stream) {
    NSP_CHECK_ARG(x && w_taps && bias && y, "conv3x3_tc: null pointer");
    NSP_CHECK_ARG(B > 0 && T > 0 && F > 0, "conv3x3_tc: bad shape");
    NSP_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)w_taps % 16 == 0) && ((uintptr_t)y % 16 == 0), "conv3x3_tc: alignment");
    void* fnp = nullptr;
    if (!get_tma_encode(&fnp)) return NSP_ERR_CUDA;
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    CUtensorMap tx, tw;
    {
        cuuint64_t gdim[4] = {32, (cuuint64_t)F, (cuuint64_t)T, (cuuint64_t)B};
        cuuint64_t gstr[3] = {64, (cuuint64_t)F * 64, (cuuint64_t)T * F * 64};
        cuuint32_t box[4] = {32, TF, TT, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3x3_tc: tensor map (x) failed: %d", (int)r); return NSP_ERR_CUDA; }
    }
    {
        cuuint64_t gdim[2] = {288, 32};
        cuuint64_t gstr[1] = {288 * 2};
        cuuint32_t box[2] = {32, 32};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_taps), gdim, gstr, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3x3_tc: tensor map (w) failed: %d", (int)r); return NSP_ERR_CUDA; }
    }
    ConvTcArgs a;
    a.bias = bias; a.y = (__nv_bfloat16*)y; a.B = B; a.T = T; a.F = F; a.pool = pool2x2; a.relu = relu;
    a.To = pool2x2 ? (T + 1) / 2 : T; a.Fo = pool2x2 ? (F + 1) / 2 : F;
    const size_t smem = 1024 + 20 * 1024 + (size_t)NST * A_BYTES + 512;
    static bool attr = false;
    if (!attr) { NSP_CUDA_OK(cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    const int tiles = B * ceil_div(T, TT) * ceil_div(F, TF);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    launch_k(conv3x3_tc_kernel, dim3(grid), dim3(192), smem, (cudaStream_t)stream, tx, tw, a);
    NSP_LAUNCH_OK();
    return NSP_OK;
}
