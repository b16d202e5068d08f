// Blackwell (sm_100a) tensor-core plumbing written as inline PTX: mbarrier, TMA (cp.async.bulk.tensor),
// TMEM allocation, tcgen05.mma / commit / ld, UMMA shared-memory and instruction descriptors.
// Bit layouts follow the PTX ISA "tcgen05" matrix-descriptor / instruction-descriptor tables
// (cross-checked against cute/arch/mma_sm100_desc.hpp in the vendored CUTLASS headers).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace nsp {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp (all 32 lanes must reach the call).  The single-thread roles (TMA producer, MMA issuer) are
// entered through this instead of `lane == 0`: ptxas knows exactly one thread is active under an elect.sync predicate and
// keeps descriptors / coordinates / barrier addresses in uniform registers; under `lane == 0` every UTCHMMA / UTMALDG is
// wrapped in a VOTEU / ELECT / R2UR / BRA.U.ANY waterfall of ~70 cycles -- more than a 128 x 32 x 16 MMA takes (measured on
// the LSTM step kernel: 36 -> 3 ns per instruction; profiles/README.md round 2).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// Same, at a point where convergence is not syntactically obvious (epilogue loops): reconverge first.  elect.sync is
// deterministic for a given member mask, so per-thread state (cp.async.bulk groups) stays with one lane across calls.
__device__ __forceinline__ bool elect_sync() { __syncwarp(); return elect_one(); }

// ---------------- mbarrier ----------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Every mbarrier wait is bounded: a lost arrival (protocol bug at an untested size class) traps instead of hanging the GPU
// (2^26 failed polls: seconds).  The retry loop is ONE inline-PTX block (block-local labels, as CUTLASS's ClusterBarrier::wait
// does): a C++ loop with a counter in it stopped nvcc from unrolling the ENCLOSING tap / k-block loops -- conv3x3_tc lost 48 %
// (1.16 -> 1.71 ms per step), the training step 0.8 ms (A/B builds, SASS 1737 vs 1153 lines; profiles/README.md round 2).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#ifdef NSP_UNBOUNDED_WAITS                       // A/B build only: the round-1 bare spin
    while (!mbar_try_wait(bar, parity)) {}
#else
    // The loop is the bare `while (!try_wait)` nvcc rotates well (fast path: one try_wait that falls through, retry loop out of
    // line -- a taken branch per wait cost conv3x3_tc 45 %); the bound lives in ONE opaque asm statement in its body.
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        asm volatile(
            "{\n\t"
            ".reg .pred q;\n\t"
            "add.u32 %0, %0, 1;\n\t"
            "setp.eq.u32 q, %0, 0x4000000;\n\t"
            "@q trap;\n\t"
            "}"
            : "+r"(polls) :: "memory");
    }
#endif
}

// Bare spin for the single TMA / MMA threads of kernels whose waits are short and many (conv3x3_tc: nine 32-wide MMAs per
// tile).  nvcc turns exactly this loop into ONE try_wait that falls through plus an out-of-line retry loop; every bounded
// form tried (counter in C++, one asm block, out-of-line call, counting asm body) left a TAKEN branch on the fast path and
// cost that kernel 45 % (1.15 -> 1.70 ms per step; A/B builds in profiles/README.md round 2).  Only for threads whose stall
// starves a BOUNDED waiter downstream (the epilogue warps' mbar_wait on the accumulator barrier): a lost arrival still traps.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

// Wait that does not spin: try_wait with a suspend-time hint parks the thread in hardware until the phase completes or the
// hint (ns) expires, so waiting warps leave the issue slots of their scheduler to the warps that have work (the CTC lattice
// warps ran 3-5x slower next to spinning producer / consumer waits: profiles/README.md round 2).
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u32 c;\n\t"
        "mov.u32 c, 0;\n"
        "NSP_WAITP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra NSP_DONEP;\n\t"
        "add.u32 c, c, 1;\n\t"
        "setp.lt.u32 p, c, 400000;\n\t"
        "@p bra NSP_WAITP;\n\t"
        "trap;\n"
        "NSP_DONEP:\n\t"
        "}"
        :: "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
}

// Epilogue-side wait (hundreds of threads waiting a whole mainloop for the accumulator).  Parking these (-DNSP_EPI_PARKED)
// made no measurable difference on the GEMM / conv kernels (A/B build, round 2), so they spin like the rest.
__device__ __forceinline__ void mbar_wait_epi(uint64_t* bar, uint32_t parity) {
#ifdef NSP_EPI_PARKED
    mbar_wait_parked(bar, parity);
#else
    mbar_wait(bar, parity);
#endif
}

// ---------------- TMA ----------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(m) : "memory");
}
// 2-D tiled load: c0 = innermost (contiguous) coordinate, c1 = row coordinate
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// ---------------- TMEM ----------------
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_holder) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(smem_holder)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {        // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tcgen05.commit: the mbarrier gets one arrival when all previously issued MMAs of this thread retire
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], bf16/fp16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// tf32 inputs (fp32 words in smem, low 13 mantissa bits ignored), fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <-> TMEM lane base+i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
// 16-column variant (fills r[0..15]; the array is declared with 32 entries for a uniform call site)
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------- shared-memory staged bulk tensor stores (epilogues) ----------------
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(m), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }

// explicit shared-space accesses (the staging pointers are derived from an aligned generic pointer, which would
// otherwise compile to generic ST.E / LD.E)
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void lds128(uint32_t addr, float& a, float& b, float& c, float& d) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr) : "memory");
}

// row r (= lane) of a 32 x 32 fp32 sub-tile, SWIZZLE_128B, tile base 1024-byte aligned
__device__ __forceinline__ void st_row_f32(uint8_t* tile, int r, const float* v) {
    const uint32_t row = smem_u32(tile) + r * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        sts128(row + ((j ^ (r & 7)) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
               __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
}
__device__ __forceinline__ void ld_row_f32(const uint8_t* tile, int r, float* v) {
    const uint32_t row = smem_u32(tile) + r * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j) lds128(row + ((j ^ (r & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
// row r of a 32 x 32 bf16 sub-tile, SWIZZLE_64B, tile base 512-byte aligned
__device__ __forceinline__ void st_row_bf16(uint8_t* tile, int r, const float* v) {
    const uint32_t row = smem_u32(tile) + r * 64;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        sts128(row + ((j ^ ((r >> 1) & 3)) << 4), pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
               pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
}

// dst[tile] += smem tile (element type and box from the tensor map; fp32 add performed by the L2)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(m), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}

// ---------------- descriptors ----------------
// K-major operand tile in shared memory, 128-byte swizzle: rows are 128 B apart, 8-row groups 1024 B apart.
// bits [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major) | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor (32-bit): [4,6) D fmt (1=F32) | [7,10) A fmt | [10,13) B fmt (0=F16 1=BF16 2=TF32)
// | bit15/16 A/B major (0 = K-major) | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc(uint32_t ab_fmt, uint32_t M, uint32_t N) {
    return (1u << 4) | (ab_fmt << 7) | (ab_fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

}  // namespace tc

// Host: build a 2-D TMA descriptor for a row-major [rows, cols] matrix with leading dimension ld (elements),
// box = [box_rows, 128 bytes] and 128-byte swizzle.  Returns false (and sets the error) on failure.
bool make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, bool is_bf16, uint64_t rows, uint64_t cols,
                  uint64_t ld, uint32_t box_rows);

// Row-major [rows, cols] matrix (pitch ld elements), box = [box_rows, box_cols]; box_cols * element size must equal the
// swizzle span.  gemm_tma_epi.cu.
bool encode_tmap_2d(CUtensorMap* out, const void* base, bool is_bf16, uint64_t rows, uint64_t cols, uint64_t ld,
                    uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swz, const char* what);

}  // namespace nsp
