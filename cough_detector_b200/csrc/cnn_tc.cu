// CoughDetectorResidual forward on the 5th-generation tensor cores (CD_PRECISION_BF16).
//
// Replaces src/model.py:253-293 (forward of the stem + two ResidualBlocks + head) in eval mode with
// BatchNorm folded into bf16 weights.  Every convolution is a GEMM whose A operand is never built:
// activations live in global memory as rows that are ALREADY shared-memory operand images (pixel = one
// swizzled K-major row), a tile of 128 output pixels is staged with one cp.async.bulk per 64-channel plane,
// and each filter tap is the same staged rows seen through a descriptor whose start address is shifted by
// whole rows (the tensor core swizzles on absolute address bits; cd_selftest_umma checks that).
//
//   producer warp 0   : weights once (resident), then one bulk copy (TMA engine) per plane per tile
//   MMA warp 1        : one elected thread issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, M = 128) into
//                       TMEM; tcgen05.commit frees operand slots / publishes accumulators
//   epilogue warps 2-5: tcgen05.ld, + folded bias, ReLU (stem: 2x2 max-pool first), store bf16 rows in the
//                       layout the next kernel stages; accumulators are double buffered in TMEM
//
//   stem_tc_kernel            7x7/s2 + ReLU + maxpool as ONE GEMM per pooled pixel over 4x4-pixel blocks
//   conv3x3_s2_tc_kernel x2   conv1 of each block (stride 2) over the phase-split input
//   conv3x3_s1_tc_kernel      conv2 of block 0 (stride 1) + the block's 1x1/s2 skip conv as a tenth tap
//   conv3x3_s1_pair_tc_kernel the same for block 1 on CTA pairs (cta_group::2, M = 256)
//   head_kernel               42-pixel average pool + Linear(128->2) on CUDA cores (fixed summation order)
#include <vector>

#include "cd_common.cuh"
#include <stdlib.h>

namespace cd {

// ----------------------------------------------------------------------------------------------
// PTX helpers (tcgen05 / mbarrier / cp.async)
// ----------------------------------------------------------------------------------------------
namespace tc {

__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (-> launch error the host reports) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();     // ~2 s at 1.9 GHz
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Programmatic dependent launch: the network's kernels are launched with programmatic stream serialization, so a
// kernel's prologue (barrier init, TMEM allocation, the resident WEIGHTS going to shared memory) overlaps the tail of the
// kernel before it.  pdl_launch_dependents: the next kernel may start being scheduled; pdl_wait: everything the
// previous kernels wrote is complete and visible (must precede the first read of an activation).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 operands, fp32 accumulate, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// EXPERIMENTAL issue form, OFF by default and not in the shipped library (`make elect`; never run on a GPU -- round 2 found the
// issue-bound profile after its GPU budget was spent, profiles/r2_mma_issue_bound.txt).  With CD_MMA_ELECT=1 the MMA warp of the
// single-CTA kernels stays CONVERGENT: all 32 lanes run the tile loop and the barrier waits, and only the tcgen05 instruction
// itself is predicated on the lane elect.sync picks (the CuTe pattern), so the compiler emits straight-line `@UP UTCHMMA` instead
// of wrapping every MMA issued from `if (lane == 0)` in an R2UR / ELECT / BRA.U.ANY retry loop.
#ifndef CD_MMA_ELECT
#define CD_MMA_ELECT 0
#endif
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(s32(bar)) : "memory");
}
#if CD_MMA_ELECT
#define CD_MMA_ISSUER(lane) true
#define CD_UMMA tc::umma_bf16_elect
#define CD_COMMIT tc::tc_commit_elect
#else
#define CD_MMA_ISSUER(lane) ((lane) == 0)
#define CD_UMMA tc::umma_bf16
#define CD_COMMIT tc::tc_commit
#endif
// ---- CTA-pair (cta_group::2) variants: one MMA spans two SMs (M = 256: 128 rows from each CTA's shared memory,
// each CTA supplies half of the N rows of B), issued by the even ("leader") CTA of a 2-CTA cluster.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (when all MMAs issued so far have retired) on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(s32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// wait on a barrier that threads of the OTHER CTA of the pair arrive on (cluster-scope acquire), bounded
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
// plain arrive on the LEADER CTA's copy of a barrier (bit 24 of a shared::cluster address selects the odd CTA of a pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(s32(bar) & 0xFEFFFFFFu) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane = tile row).  The _issue form does
// not wait: several loads can be in flight under ONE tmem_ld_wait (an epilogue that waits after every load pays the full
// TMEM latency once per 32 columns).
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    tmem_ld32_issue(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major operand tile in the canonical SWIZZLE_128B layout (cute Layout_K_SW128_Atom): rows of
// 64 bf16 = 128 B, 16-byte chunk j of row r stored at chunk j ^ (r & 7), 8-row groups 1024 B
// apart.  Descriptor fields per cute/arch/mma_sm100_desc.hpp (SmemDescriptor): start>>4 [0,14),
// LBO>>4 [16,30) (=1, unused for swizzled K-major), SBO>>4 [32,46) (=64), version=1 [46,48),
// layout_type=2 (SWIZZLE_128B) [61,64).  Tile base must be 1024-byte aligned.
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)64 << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// InstrDescriptor (same header): c_format F32=1 [4,6), a/b_format BF16=1 [7,10)/[10,13), both
// K-major (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Advance a descriptor's start address by `bytes` (multiple of 16; the 14-bit field holds addr >> 4 and
// shared memory is < 256 KB, so the add never carries out of the field).  One 64-bit add per MMA operand
// instead of rebuilding the descriptor keeps the single issuing thread off the critical path.
__device__ __forceinline__ uint64_t desc_adv(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }
// SWIZZLE_32B K-major: rows of 16 bf16 = 32 B (exactly one K16 step), 8-row groups 256 B apart, 16-byte chunk c
// of row r at c ^ ((r >> 2) & 1) (Swizzle<1,4,3>: address bit 4 ^= bit 7).  layout_type = 6, SBO = 256 B.
__device__ __forceinline__ uint64_t desc_sw32(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)16 << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}
__device__ __forceinline__ uint32_t swz32(int row, int chunk) { return (uint32_t)(row * 32 + ((chunk ^ ((row >> 2) & 1)) << 4)); }
__device__ __forceinline__ uint32_t swz128(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

}  // namespace tc

// ----------------------------------------------------------------------------------------------
// UMMA self-test: D[128 x N] = A[128 x K] * B[N x K]^T through the exact helpers the conv kernels use
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
umma_selftest_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Bm, float* __restrict__ D,
                     int N, int K, int shift, int mode) {
    extern __shared__ unsigned char raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    const int kc_n = K / 64;
    const int rowsA = (128 + shift + 7) & ~7;             // A has 128 + shift rows; D[i] = A[i + shift] . B^T
    const int a_chunk = rowsA * 128;
    unsigned char* sA = base;                             // [kc][rowsA][128 B]
    unsigned char* sB = base + (size_t)kc_n * a_chunk;    // [kc][N rows][128 B]
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + shift) * (K / 8); i += 128) {
        const int row = i / (K / 8), ch = i % (K / 8), kc = ch >> 3, j = ch & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)row * K + ch * 8);
        if (mode == 2)      // sub-tile per K16 step: [rowsA][32 B], 16-byte chunk c of row r at c ^ ((r >> 2) & 1)
            *reinterpret_cast<uint4*>(sA + (ch >> 1) * (rowsA * 32) + tc::swz32(row, ch & 1)) = v;
        else
            *reinterpret_cast<uint4*>(sA + kc * a_chunk + tc::swz128(row, j)) = v;
    }
    for (int i = tid; i < N * (K / 8); i += 128) {
        const int row = i / (K / 8), ch = i % (K / 8), kc = ch >> 3, j = ch & 7;
        *reinterpret_cast<uint4*>(sB + (size_t)kc * N * 128 + tc::swz128(row, j)) = *reinterpret_cast<const uint4*>(Bm + (size_t)row * K + ch * 8);
    }
    tc::fence_proxy_async();
    if (warp == 0) tc::tmem_alloc(&tmem_base, 256);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_bf16(128, N);
        for (int kc = 0; kc < kc_n; ++kc)
            for (int k4 = 0; k4 < 4; ++k4) {
                // shifted view: start the A operand `shift` rows (128 B each) into the tile
                uint64_t ad = tc::desc_sw128(tc::s32(sA + kc * a_chunk) + shift * 128 + k4 * 32);
                if (mode == 1) ad |= (uint64_t)(shift & 7) << 49;      // base_offset variant
                if (mode == 2) ad = tc::desc_sw32(tc::s32(sA + (kc * 4 + k4) * (rowsA * 32)) + shift * 32);
                tc::umma_bf16(tm, ad, tc::desc_sw128(tc::s32(sB + (size_t)kc * N * 128) + k4 * 32), idesc, (kc | k4) != 0);
            }
        tc::tc_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 32; ++i) D[(size_t)tid * N + c0 + i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 256);
}

// CTA-pair self-test: D[256 x N] = A[256 x K] * B[N x K]^T with one cta_group::2 MMA stream; CTA r of the
// cluster holds A rows [128 r, 128 r + 128) and B rows [N/2 r, N/2 r + N/2), and reads back its 128 rows of D.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
umma_pair_selftest_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Bm, float* __restrict__ D,
                          int N, int K) {
    extern __shared__ unsigned char raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    const int kc_n = K / 64, nh = N / 2;
    unsigned char* sA = base;                             // [kc][128 rows][128 B]
    unsigned char* sB = base + (size_t)kc_n * 16384;      // [kc][N/2 rows][128 B]
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = tc::cluster_ctarank();
    for (int i = tid; i < 128 * (K / 8); i += 128) {
        const int row = i / (K / 8), ch = i % (K / 8), kc = ch >> 3, j = ch & 7;
        *reinterpret_cast<uint4*>(sA + kc * 16384 + tc::swz128(row, j)) =
            *reinterpret_cast<const uint4*>(A + (size_t)(rank * 128 + row) * K + ch * 8);
    }
    for (int i = tid; i < nh * (K / 8); i += 128) {
        const int row = i / (K / 8), ch = i % (K / 8), kc = ch >> 3, j = ch & 7;
        *reinterpret_cast<uint4*>(sB + (size_t)kc * nh * 128 + tc::swz128(row, j)) =
            *reinterpret_cast<const uint4*>(Bm + (size_t)(rank * nh + row) * K + ch * 8);
    }
    tc::fence_proxy_async();
    if (warp == 0) tc::tmem_alloc2(&tmem_base, 256);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync();                                   // both CTAs' operands, barriers and TMEM are ready
    tc::tc_fence_after();
    const uint32_t tm = tmem_base;
    if (rank == 0 && tid == 0) {
        const uint32_t idesc = tc::idesc_bf16(256, N);
        for (int kc = 0; kc < kc_n; ++kc)
            for (int k4 = 0; k4 < 4; ++k4)
                tc::umma2_bf16(tm, tc::desc_sw128(tc::s32(sA + kc * 16384) + k4 * 32),
                               tc::desc_sw128(tc::s32(sB + (size_t)kc * nh * 128) + k4 * 32), idesc, (kc | k4) != 0);
        tc::tc_commit2(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 32; ++i) D[(size_t)(rank * 128 + tid) * N + c0 + i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync();
    if (warp == 0) tc::tmem_dealloc2(tm, 256);
}

// ----------------------------------------------------------------------------------------------
// activation layouts of the convolution kernels
// ----------------------------------------------------------------------------------------------
// Padded + swizzled ("PS") activation layout consumed by the stride-1 kernel: pixels are 128-byte rows of 64
// channels (C = 128: two such planes) on a grid of pitch WP = W + 1 with H + 1 rows per clip; global row
// rg = clip * RPC + (y+1) * WP + (x+1); the 16-byte chunk j of a row is stored at chunk j ^ (rg & 7), i.e.
// already in the SWIZZLE_128B pattern of a tile that starts at a multiple of 8 rows.  With that, the A
// operand of tap (ky, kx) for 128 consecutive output rows m' = y * WP + x is the SAME staged rows viewed
// (ky * WP + kx) rows further down -- no im2col copy.  The zero padding is SHARED: column 0 of a grid row is the
// left pad of that row and, being the next index after column W of the row above, also the right pad of that
// one; grid row 0 of a clip is its top pad and the bottom pad of the clip before it.  So a clip costs
// (H+1)(W+1) GEMM rows instead of (H+2)(W+2) (block 0: 168 vs 195, block 1: 56 vs 72).  Pad rows are zero
// because the buffer is zeroed once and only interior pixels are ever written.
template <int H, int W>
struct PsLayout {
    static constexpr int HP = H + 1, WP = W + 1, RPC = HP * WP;
    static constexpr int ROWS_IN = (128 + 2 * WP + 2 + 7) & ~7;     // rows one 128-row tile reads
};

// Phase-split variant for the INPUT of a stride-2 convolution (see conv3x3_s2_tc_kernel): phase (py, px) holds the
// padded pixels (2a + py, 2b + px).  Taps read (oy + (ky >> 1), ox + (kx >> 1)), i.e. offset +1 only in the even
// phases, whose element 0 is padding -- so here too the last column / row wraps onto the next row's / clip's
// pad: A2 = (H + 2) / 2 rows of pitch B2 = (W + 2) / 2 (22 x 25 -> 12 x 13, 11 x 13 -> 6 x 7 = exactly the outputs).
template <int H, int W>
struct Ps2Layout {
    static constexpr int A2 = (H + 2) / 2, B2 = (W + 2) / 2, RPC = A2 * B2;
    static constexpr int ROWS_IN = (128 + B2 + 1 + 7) & ~7;
};

// ----------------------------------------------------------------------------------------------
// stride-1 3x3 convolution without im2col: shifted descriptor views over PS-layout rows
// ----------------------------------------------------------------------------------------------
// conv2 of each residual block (C -> C, pad 1) + bias + identity add + ReLU.  GEMM rows are the padded
// pixel rows m' of PsLayout (valid outputs: x < W, y < H; the rest are computed and dropped, 15 % / 25 %
// of the MMA rows).  Per 128-row tile ONE bulk copy per
// 64-channel plane stages ROWS_IN rows (20 KB) instead of nine im2col chunks (144 KB), and the weights
// (9 taps x C/64 planes x [64 x 128 B]) stay resident in shared memory: a CTA owns one 64-wide half of
// the output channels (C = 128: two halves -> the A rows are read twice, the 288 KB of weights never).
constexpr int kS1Threads = 192;                   // warp 0: bulk-copy producer, warp 1: MMA, warps 2-5: epilogue

// The block's 1x1/s2 skip projection rides along as a TENTH tap: its operand (the block input sampled at
// (2y, 2x), written in PS layout by whoever produced that input) is one more 64-channel plane, viewed at
// the centre shift WP + 1, multiplied by the skip weights into the SAME accumulator -- the identity path
// never leaves TMEM and no fp32 identity tensor exists.
template <int C, int H, int W>
struct S1Shape {
    using PS = PsLayout<H, W>;
    static constexpr int NPL = C / 64;                            // 64-channel planes of the conv input
    static constexpr int NPLT = NPL + 1;                          // + the skip-operand plane
    static constexpr int HALVES = C / 64;                         // 64-wide halves of the output channels
    static constexpr int A_BYTES = PS::ROWS_IN * 128;             // one staged plane
    static constexpr int STAGES = C == 64 ? 3 : 1;                // tiles in flight; every plane has its own barriers,
                                                                  // so even STAGES = 1 refills plane q while q+1 computes
    static constexpr int NCHUNK = 9 * NPL + 1;                    // resident weight chunks of one half (last = skip)
    static constexpr int W_BYTES = NCHUNK * 8192;
    static constexpr size_t SMEM = 1024 + (size_t)STAGES * NPLT * A_BYTES + W_BYTES + 64 * 4 + 64;
};

struct S1Args {
    const __nv_bfloat16* x;        // PS layout: NPL conv planes then the skip plane, ps_plane_elems apart
    const __nv_bfloat16* w;        // packed conv chunks [9 * NPL][C rows x 128 B]
    const __nv_bfloat16* wskip;    // packed skip chunk [C rows x 128 B]
    const float* bias;             // [C] conv2 bias + skip bias
    __nv_bfloat16* y_bf16;         // [n * H * W][C] bf16 (OUT_F32 == false)
    float* y_f32;                  // [n * H * W][C] fp32 (single-CTA kernel, OUT_F32)
    float* psum;                   // pair kernel: [n][H][128] fp32, sums of the block output over the W pixels of an image row
    __nv_bfloat16* xs_out;         // WRITE_XS: next block's skip plane (PS layout of the (H+1)/2 x (W+1)/2 grid)
    int64_t y_plane_elems;         // WRITE_XS: y_bf16 is written phase-split (Ps2Layout<H, W>, 4 planes this far apart)
    int64_t ps_plane_elems;
    int n;                         // clips
    int rev;                       // walk the tiles from the last to the first (see tile_order)
};

template <int C, int H, int W, bool OUT_F32, bool WRITE_XS>
__global__ void __launch_bounds__(kS1Threads, 1)
conv3x3_s1_tc_kernel(const S1Args a) {
    const bool REV_ORDER = a.rev != 0;
    using S = S1Shape<C, H, W>;
    using PS = PsLayout<H, W>;
    using PSN = PsLayout<(H + 1) / 2, (W + 1) / 2>;               // the next block's grid (stride 2, pad 1, k 3)
    constexpr int NPOS = H * W;
    constexpr uint32_t IDESC = tc::idesc_bf16(128, 64);
    constexpr int TMEM_COLS = 128;                                // 2 x 64 fp32 columns
    constexpr int NB = S::STAGES * S::NPLT;

    extern __shared__ unsigned char raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    unsigned char* s_w = base + (size_t)NB * S::A_BYTES;
    float* s_bias = reinterpret_cast<float*>(s_w + S::W_BYTES);
    __shared__ uint64_t full_bar[NB], empty_bar[NB], tfull_bar[2], tempty_bar[2], w_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int half = blockIdx.x % S::HALVES;                      // which 64 output channels this CTA owns
    const int cta = blockIdx.x / S::HALVES, n_cta = gridDim.x / S::HALVES;
    const int64_t rows = (int64_t)a.n * PS::RPC;
    const int n_tiles = (int)((rows + 127) / 128);

    if (tid == 0) {
        for (int i = 0; i < NB; ++i) {
            tc::mbar_init(&full_bar[i], 1);
            tc::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&tfull_bar[i], 1);
            tc::mbar_init(&tempty_bar[i], 128);
        }
        tc::mbar_init(&w_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 64) s_bias[tid] = a.bias[half * 64 + tid];
    if (warp == 1) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    tc::pdl_launch_dependents();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ======================= producer: weights once, then one bulk copy per plane per tile =========
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(&w_bar, S::W_BYTES);
            for (int q = 0; q < 9 * S::NPL; ++q)      // rows [64 half, 64 half + 64) of every packed chunk
                tc::bulk_g2s(s_w + q * 8192, a.w + (size_t)q * C * 64 + (size_t)half * 64 * 64, 8192, &w_bar);
            tc::bulk_g2s(s_w + 9 * S::NPL * 8192, a.wskip + (size_t)half * 64 * 64, 8192, &w_bar);
            tc::pdl_wait();                            // weights are in flight; the activations need the previous kernel
            int it = 0;
            for (int t_ = cta; t_ < n_tiles; t_ += n_cta, ++it) {
                [[maybe_unused]] const int tile = REV_ORDER ? n_tiles - 1 - t_ : t_;
                const int s = it % S::STAGES;
                const uint32_t par = (uint32_t)(((it / S::STAGES) & 1) ^ 1);
#pragma unroll
                for (int q = 0; q < S::NPLT; ++q) {
                    const int bi = s * S::NPLT + q;
                    tc::mbar_wait(&empty_bar[bi], par);
                    tc::mbar_arrive_expect_tx(&full_bar[bi], S::A_BYTES);
                    tc::bulk_g2s(base + (size_t)bi * S::A_BYTES, a.x + q * a.ps_plane_elems + (int64_t)tile * 128 * 64,
                                 S::A_BYTES, &full_bar[bi]);
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer: per plane 9 taps x 4 K-steps, skip plane 1 tap ===========
        if (CD_MMA_ISSUER(lane)) {
            tc::mbar_wait(&w_bar, 0);
            int it = 0;
            for (int t_ = cta; t_ < n_tiles; t_ += n_cta, ++it) {
                [[maybe_unused]] const int tile = REV_ORDER ? n_tiles - 1 - t_ : t_;
                const int s = it % S::STAGES, acc = it & 1;
                const uint32_t par = (uint32_t)((it / S::STAGES) & 1);
                tc::mbar_wait(&tempty_bar[acc], (uint32_t)(((it >> 1) & 1) ^ 1));
                const uint32_t d = tmem_base + acc * 64;
                const uint64_t dw = tc::desc_sw128(tc::s32(s_w));
#pragma unroll
                for (int q = 0; q < S::NPLT; ++q) {
                    const int bi = s * S::NPLT + q;
                    tc::mbar_wait(&full_bar[bi], par);
                    tc::tc_fence_after();
                    const uint64_t da = tc::desc_sw128(tc::s32(base + (size_t)bi * S::A_BYTES));
                    if (q < S::NPL) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint32_t shift = (uint32_t)((tap / 3) * PS::WP + (tap % 3)) * 128;   // rows down the image
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                CD_UMMA(d, tc::desc_adv(da, shift + k4 * 32),
                                              tc::desc_adv(dw, (tap * S::NPL + q) * 8192 + k4 * 32), IDESC,
                                              (q | tap | k4) != 0);
                        }
                    } else {
                        const uint32_t shift = (uint32_t)(PS::WP + 1) * 128;                          // centre tap
                        // the skip projection reads the block's INPUT channels (C / 2): with C = 64 only the first 32 of a
                        // skip-plane row's 64 channel slots are populated (the rest, and their weights, are zero)
                        constexpr int kSkipSteps = C == 64 ? 2 : 4;
#pragma unroll
                        for (int k4 = 0; k4 < kSkipSteps; ++k4)
                            CD_UMMA(d, tc::desc_adv(da, shift + k4 * 32),
                                          tc::desc_adv(dw, 9 * S::NPL * 8192 + k4 * 32), IDESC, 1u);
                    }
                    CD_COMMIT(&empty_bar[bi]);        // this plane's slot can be refilled for the next tile
                }
                CD_COMMIT(&tfull_bar[acc]);
            }
        }
    } else {
        // ======================= epilogue: + bias, ReLU, store the valid pixels ======================
        const int lane_grp = warp & 3;
        int it = 0;
        for (int t_ = cta; t_ < n_tiles; t_ += n_cta, ++it) {
                [[maybe_unused]] const int tile = REV_ORDER ? n_tiles - 1 - t_ : t_;
            const int acc = it & 1;
            const int64_t mp = (int64_t)tile * 128 + lane_grp * 32 + lane;      // padded-row index m'
            const int clip = (int)(mp / PS::RPC), rem = (int)(mp - (int64_t)clip * PS::RPC);
            const int y = rem / PS::WP, x = rem - y * PS::WP;
            const bool live = clip < a.n && y < H && x < W;
            const int64_t m = (int64_t)clip * NPOS + y * W + x;
            tc::mbar_wait(&tfull_bar[acc], (uint32_t)((it >> 1) & 1));
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + acc * 64;
            float v0[32], v1[32];
            {
                uint32_t r0[32], r1[32];
                tc::tmem_ld32_issue(taddr, r0);
                tc::tmem_ld32_issue(taddr + 32, r1);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) { v0[i] = __uint_as_float(r0[i]); v1[i] = __uint_as_float(r1[i]); }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(&tempty_bar[acc]);
            if (live) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    v0[i] = fmaxf(v0[i] + s_bias[i], 0.f);
                    v1[i] = fmaxf(v1[i] + s_bias[32 + i], 0.f);
                }
                if (OUT_F32) {
                    float4* op = reinterpret_cast<float4*>(a.y_f32 + m * C + half * 64);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        op[q] = make_float4(v0[4 * q], v0[4 * q + 1], v0[4 * q + 2], v0[4 * q + 3]);
                        op[8 + q] = make_float4(v1[4 * q], v1[4 * q + 1], v1[4 * q + 2], v1[4 * q + 3]);
                    }
                } else {
                    // WRITE_XS (block 0): the activation feeds the next block's stride-2 conv, so it is stored
                    // phase-split (Ps2Layout<H, W>); otherwise plain NHWC rows
                    using P2 = Ps2Layout<H, W>;
                    const int64_t r2 = (int64_t)clip * P2::RPC + ((y + 1) >> 1) * P2::B2 + ((x + 1) >> 1);
                    __nv_bfloat16* op2 = a.y_bf16 + (int64_t)(((y + 1) & 1) * 2 + ((x + 1) & 1)) * a.y_plane_elems + r2 * 64;
                    uint4* op = reinterpret_cast<uint4*>(a.y_bf16 + m * C + half * 64);
                    // the next block's skip conv reads this activation at (2y', 2x'): emit that sample as a PS row
                    const bool xs = WRITE_XS && !(y & 1) && !(x & 1);
                    const int64_t rg2 = (int64_t)clip * PSN::RPC + (y / 2 + 1) * PSN::WP + (x / 2 + 1);
                    __nv_bfloat16* xrow = a.xs_out + rg2 * 64;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float* v = q < 4 ? v0 + 8 * q : v1 + 8 * (q - 4);
                        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
                        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
                        const uint4 pk = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                                                    *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
                        if (WRITE_XS) *reinterpret_cast<uint4*>(op2 + ((q ^ (int)(r2 & 7)) << 3)) = pk;
                        else op[q] = pk;
                        if (xs) *reinterpret_cast<uint4*>(xrow + ((q ^ (int)(rg2 & 7)) << 3)) = pk;
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ----------------------------------------------------------------------------------------------
// stride-1 convolution of block 1 (C = 128) on CTA PAIRS (cta_group::2)
// ----------------------------------------------------------------------------------------------
// The single-CTA kernel above keeps only one 64-channel half of the 288 KB of weights resident, so every tile
// is staged and multiplied twice.  A 2-CTA cluster does it once: CTA r of the pair stages ITS OWN tile (128 rows)
// and holds weight rows [64 r, 64 r + 64) of every chunk; one tcgen05.mma.cta_group::2 (M = 256, N = 128), issued
// by the even CTA, multiplies both tiles by all 128 output channels, reading A from each CTA's shared memory and
// half of B from each.  Per 128 x 128 x 16 MACs a CTA fetches 4 KB of A + 2 KB of B instead of 2 x (4 + 2) KB --
// the tensor pipe here is bound by operand fetch from shared memory (about 64 B/clk/SM measured), not by math.
//   leader (rank 0) warp 1: MMA issuer; waits its own slots AND the peer's (peer warp 1 relays its slot-full
//   barriers with a remote arrive); tcgen05.commit multicasts "slot free" / "accumulator ready" to both CTAs;
//   both epilogues hand accumulators back by arriving on the leader's barrier.
template <int H, int W>
struct S1PairShape {
    using PS = PsLayout<H, W>;
    static constexpr int NPLT = 3;                                // two conv planes + the skip-operand plane
    static constexpr int A_BYTES = PS::ROWS_IN * 128;
    static constexpr int NCHUNK = 19;                             // 9 taps x 2 planes + skip, 64 weight rows each
    static constexpr int W_BYTES = NCHUNK * 8192;
    static constexpr size_t SMEM = 1024 + (size_t)NPLT * A_BYTES + W_BYTES + 128 * 4 + 64;
};

template <int H, int W>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kS1Threads, 1)
conv3x3_s1_pair_tc_kernel(const S1Args a) {
    const bool REV_ORDER = a.rev != 0;
    constexpr int C = 128;
    using S = S1PairShape<H, W>;
    using PS = PsLayout<H, W>;
    constexpr int NPOS = H * W;
    constexpr uint32_t IDESC = tc::idesc_bf16(256, 128);
    constexpr int TMEM_COLS = 256;                                // 2 accumulators x 128 fp32 columns
    constexpr int NB = S::NPLT;

    extern __shared__ unsigned char raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    unsigned char* s_w = base + (size_t)NB * S::A_BYTES;
    float* s_bias = reinterpret_cast<float*>(s_w + S::W_BYTES);
    __shared__ uint64_t full_bar[NB], peer_full_bar[NB], empty_bar[NB], tfull_bar[2], tempty_bar[2], w_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int64_t rows = (int64_t)a.n * PS::RPC;
    const int n_tiles = (int)((rows + 127) / 128);
    const int n_tp = (n_tiles + 1) / 2;                           // tile pairs; CTA r takes tile 2 tp + r

    if (tid == 0) {
        for (int i = 0; i < NB; ++i) {
            tc::mbar_init(&full_bar[i], 1);
            tc::mbar_init(&peer_full_bar[i], 1);
            tc::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&tfull_bar[i], 1);
            tc::mbar_init(&tempty_bar[i], 256);                   // the epilogue threads of BOTH CTAs
        }
        tc::mbar_init(&w_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 128) s_bias[tid] = a.bias[tid];
    if (warp == 1) tc::tmem_alloc2(&tmem_base_s, TMEM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync();
    tc::tc_fence_after();
    tc::pdl_launch_dependents();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ======================= producer (both CTAs): my weight half once, my tile's planes ==========
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(&w_bar, S::W_BYTES);
            for (int q = 0; q < 18; ++q)
                tc::bulk_g2s(s_w + q * 8192, a.w + (size_t)q * C * 64 + (size_t)rank * 64 * 64, 8192, &w_bar);
            tc::bulk_g2s(s_w + 18 * 8192, a.wskip + (size_t)rank * 64 * 64, 8192, &w_bar);
            tc::pdl_wait();
            int it = 0;
            for (int t_ = pair; t_ < n_tp; t_ += n_pairs, ++it) {
                [[maybe_unused]] const int tp = REV_ORDER ? n_tp - 1 - t_ : t_;
                const int64_t tile = 2 * (int64_t)tp + rank;
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    tc::mbar_wait(&empty_bar[q], (uint32_t)((it & 1) ^ 1));
                    tc::mbar_arrive_expect_tx(&full_bar[q], S::A_BYTES);
                    tc::bulk_g2s(base + (size_t)q * S::A_BYTES, a.x + q * a.ps_plane_elems + tile * 128 * 64, S::A_BYTES,
                                 &full_bar[q]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            tc::mbar_wait(&w_bar, 0);
            if (rank == 0) {
                // ======================= MMA issuer (leader) ==============================================
                const uint64_t dw = tc::desc_sw128(tc::s32(s_w));
                int it = 0;
                for (int t_ = pair; t_ < n_tp; t_ += n_pairs, ++it) {
                [[maybe_unused]] const int tp = REV_ORDER ? n_tp - 1 - t_ : t_;
                    const int acc = it & 1;
                    tc::mbar_wait_cluster(&tempty_bar[acc], (uint32_t)(((it >> 1) & 1) ^ 1));
                    const uint32_t d = tmem_base + acc * 128;
#pragma unroll
                    for (int q = 0; q < NB; ++q) {
                        tc::mbar_wait(&full_bar[q], (uint32_t)(it & 1));
                        tc::mbar_wait_cluster(&peer_full_bar[q], (uint32_t)(it & 1));
                        tc::tc_fence_after();
                        const uint64_t da = tc::desc_sw128(tc::s32(base + (size_t)q * S::A_BYTES));
                        if (q < 2) {
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const uint32_t shift = (uint32_t)((tap / 3) * PS::WP + (tap % 3)) * 128;
#pragma unroll
                                for (int k4 = 0; k4 < 4; ++k4)
                                    tc::umma2_bf16(d, tc::desc_adv(da, shift + k4 * 32),
                                                   tc::desc_adv(dw, (tap * 2 + q) * 8192 + k4 * 32), IDESC, (q | tap | k4) != 0);
                            }
                        } else {
                            const uint32_t shift = (uint32_t)(PS::WP + 1) * 128;
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                tc::umma2_bf16(d, tc::desc_adv(da, shift + k4 * 32), tc::desc_adv(dw, 18 * 8192 + k4 * 32), IDESC, 1u);
                        }
                        tc::tc_commit2(&empty_bar[q]);           // slot q is free again in BOTH CTAs
                    }
                    tc::tc_commit2(&tfull_bar[acc]);
                }
            } else {
                // ======================= relay (peer): tell the leader when my slots are full ============
                int it = 0;
                for (int t_ = pair; t_ < n_tp; t_ += n_pairs, ++it) {
                [[maybe_unused]] const int tp = REV_ORDER ? n_tp - 1 - t_ : t_;
#pragma unroll
                    for (int q = 0; q < NB; ++q) {
                        tc::mbar_wait(&full_bar[q], (uint32_t)(it & 1));
                        tc::mbar_arrive_leader(&peer_full_bar[q]);
                    }
                }
            }
        }
    } else {
        // ======================= epilogue (both CTAs): my 128 rows x all 128 channels ================
        const int lane_grp = warp & 3;
        int it = 0;
        for (int t_ = pair; t_ < n_tp; t_ += n_pairs, ++it) {
                [[maybe_unused]] const int tp = REV_ORDER ? n_tp - 1 - t_ : t_;
            const int acc = it & 1;
            const int64_t mp = (2 * (int64_t)tp + rank) * 128 + lane_grp * 32 + lane;
            const int clip = (int)(mp / PS::RPC), rem = (int)(mp - (int64_t)clip * PS::RPC);
            const int y = rem / PS::WP, x = rem - y * PS::WP;
            const bool live = clip < a.n && y < H && x < W;
            tc::mbar_wait(&tfull_bar[acc], (uint32_t)((it >> 1) & 1));
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + acc * 128;
            // The global average pool starts here: the eight lanes of a group hold the eight grid columns of ONE image
            // row (clip rows are multiples of 8 GEMM rows, tiles multiples of 128), so a 3-step transpose-reduce over
            // the group leaves every lane with the row sums of 4 of the 32 channels -- a fixed tree over the pixel
            // positions, independent of where the clip sits in the batch.  Only [n][H][128] row sums leave the SM
            // (3 KB per clip instead of the 21.5 KB fp32 activation); head_kernel adds the H rows and applies fc.
            const bool up4 = lane & 4, up2 = lane & 2, up1 = lane & 1;
            const int cbase = (up4 ? 16 : 0) + (up2 ? 8 : 0) + (up1 ? 4 : 0);
            float* prow = a.psum + ((int64_t)clip * H + y) * C + cbase;
            const bool row_live = clip < a.n && y < H;
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                float v[32];
                tc::tmem_ld32(taddr + c0, v);
                if (c0 == 96) {                                   // accumulator drained: hand it back to the leader
                    tc::tc_fence_before();
                    tc::mbar_arrive_leader(&tempty_bar[acc]);
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = live ? fmaxf(v[i] + s_bias[c0 + i], 0.f) : 0.f;
                float r16[16], r8[8], r4[4];
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    r16[i] = (up4 ? v[i + 16] : v[i]) + __shfl_xor_sync(0xffffffffu, up4 ? v[i] : v[i + 16], 4);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    r8[i] = (up2 ? r16[i + 8] : r16[i]) + __shfl_xor_sync(0xffffffffu, up2 ? r16[i] : r16[i + 8], 2);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    r4[i] = (up1 ? r8[i + 4] : r8[i]) + __shfl_xor_sync(0xffffffffu, up1 ? r8[i] : r8[i + 4], 1);
                if (row_live) *reinterpret_cast<float4*>(prow + c0) = make_float4(r4[0], r4[1], r4[2], r4[3]);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync();
    if (warp == 1) tc::tmem_dealloc2(tmem_base, TMEM_COLS);
}

// ----------------------------------------------------------------------------------------------
// stride-2 3x3 convolution without im2col: phase-split PS input ("PS2") + shifted descriptor views
// ----------------------------------------------------------------------------------------------
// With padded coordinates iy' = iy + 1, ix' = ix + 1 the tap (ky, kx) of output (oy, ox) reads
// iy' = 2 oy + ky, ix' = 2 ox + kx.  Splitting the padded image by the parities (iy' & 1, ix' & 1) gives
// four phase images of ceil((H+2)/2) x ceil((W+2)/2) pixels in which that tap is the pixel
// (oy + (ky >> 1), ox + (kx >> 1)) of phase (ky & 1, kx & 1): a unit-stride shifted view again.
// Rows of a phase image are 128 bytes: C = 64 -> one phase per plane (4 planes); C = 32 -> the two ix'
// parities of one iy' parity share a row (2 planes, the kx parity selects the 64-byte half = K offset).
template <int CIN, int COUT, int H, int W>
struct S2Shape {
    using P2 = Ps2Layout<H, W>;
    static constexpr int NPL = CIN == 32 ? 2 : 4;
    static constexpr int A_BYTES = P2::ROWS_IN * 128;
    static constexpr int NCH = (9 * CIN + 63) / 64;               // packed weight chunks [COUT rows x 128 B]
    static constexpr int W_BYTES = NCH * COUT * 128;
    static constexpr int STAGES = CIN == 32 ? 4 : 1;              // per-plane barriers keep even 1 stage overlapped
    static constexpr size_t SMEM = 1024 + (size_t)STAGES * NPL * A_BYTES + W_BYTES + COUT * 4 + 64;
};

struct S2Args {
    const __nv_bfloat16* x;        // PS2 planes, in_plane_elems apart
    const __nv_bfloat16* w;        // packed chunks [NCH][COUT rows x 128 B]
    const float* bias;             // [COUT]
    __nv_bfloat16* y;              // PS layout of the (HO, WO) output, COUT/64 planes out_plane_elems apart
    int64_t in_plane_elems, out_plane_elems;
    int n;
    int rev;                       // walk the tiles from the last to the first (see tile_order)
};

template <int CIN, int COUT, int H, int W>
__global__ void __launch_bounds__(kS1Threads, 1)
conv3x3_s2_tc_kernel(const S2Args a) {
    const bool REV_ORDER = a.rev != 0;
    using S = S2Shape<CIN, COUT, H, W>;
    using P2 = Ps2Layout<H, W>;
    constexpr int HO = (H - 1) / 2 + 1, WO = (W - 1) / 2 + 1;
    using PSO = PsLayout<HO, WO>;
    constexpr uint32_t IDESC = tc::idesc_bf16(128, COUT);
    constexpr int TMEM_COLS = 2 * COUT <= 128 ? 128 : 256;
    constexpr int NB = S::STAGES * S::NPL;

    extern __shared__ unsigned char raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    unsigned char* s_w = base + (size_t)NB * S::A_BYTES;
    float* s_bias = reinterpret_cast<float*>(s_w + S::W_BYTES);
    __shared__ uint64_t full_bar[NB], empty_bar[NB], tfull_bar[2], tempty_bar[2], w_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t rows = (int64_t)a.n * P2::RPC;
    const int n_tiles = (int)((rows + 127) / 128);

    if (tid == 0) {
        for (int i = 0; i < NB; ++i) {
            tc::mbar_init(&full_bar[i], 1);
            tc::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&tfull_bar[i], 1);
            tc::mbar_init(&tempty_bar[i], 128);
        }
        tc::mbar_init(&w_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < COUT) s_bias[tid] = a.bias[tid];
    if (warp == 1) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    tc::pdl_launch_dependents();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ======================= producer ============================================================
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(&w_bar, S::W_BYTES);
            for (int q = 0; q < S::NCH; ++q)
                tc::bulk_g2s(s_w + (size_t)q * COUT * 128, a.w + (size_t)q * COUT * 64, COUT * 128, &w_bar);
            tc::pdl_wait();
            int it = 0;
            for (int t_ = blockIdx.x; t_ < n_tiles; t_ += gridDim.x, ++it) {
                [[maybe_unused]] const int tile = REV_ORDER ? n_tiles - 1 - t_ : t_;
                const int s = it % S::STAGES;
                const uint32_t par = (uint32_t)(((it / S::STAGES) & 1) ^ 1);
#pragma unroll
                for (int q = 0; q < S::NPL; ++q) {
                    const int bi = s * S::NPL + q;
                    tc::mbar_wait(&empty_bar[bi], par);
                    tc::mbar_arrive_expect_tx(&full_bar[bi], S::A_BYTES);
                    tc::bulk_g2s(base + (size_t)bi * S::A_BYTES, a.x + q * a.in_plane_elems + (int64_t)tile * 128 * 64,
                                 S::A_BYTES, &full_bar[bi]);
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer: taps grouped by the plane (phase) they read =============
        if (CD_MMA_ISSUER(lane)) {
            tc::mbar_wait(&w_bar, 0);
            int it = 0;
            for (int t_ = blockIdx.x; t_ < n_tiles; t_ += gridDim.x, ++it) {
                [[maybe_unused]] const int tile = REV_ORDER ? n_tiles - 1 - t_ : t_;
                const int s = it % S::STAGES, acc = it & 1;
                const uint32_t par = (uint32_t)((it / S::STAGES) & 1);
                tc::mbar_wait(&tempty_bar[acc], (uint32_t)(((it >> 1) & 1) ^ 1));
                const uint32_t d = tmem_base + acc * COUT;
                const uint64_t dw = tc::desc_sw128(tc::s32(s_w));
                uint32_t accum = 0;
#pragma unroll
                for (int q = 0; q < S::NPL; ++q) {
                    const int bi = s * S::NPL + q;
                    tc::mbar_wait(&full_bar[bi], par);
                    tc::tc_fence_after();
                    const uint64_t da = tc::desc_sw128(tc::s32(base + (size_t)bi * S::A_BYTES));
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int ky = tap / 3, kx = tap % 3;
                        const int plane = CIN == 32 ? (ky & 1) : (ky & 1) * 2 + (kx & 1);
                        if (plane != q) continue;
                        const uint32_t shift = (uint32_t)((ky >> 1) * P2::B2 + (kx >> 1)) * 128;
                        if (CIN == 32) {      // 32 channels = 2 K-steps in the (kx & 1) half of the row
#pragma unroll
                            for (int hk = 0; hk < 2; ++hk) {
                                CD_UMMA(d, tc::desc_adv(da, shift + (kx & 1) * 64 + hk * 32),
                                              tc::desc_adv(dw, (tap >> 1) * COUT * 128 + ((tap & 1) * 2 + hk) * 32), IDESC, accum);
                                accum = 1;
                            }
                        } else {
#pragma unroll
                            for (int hk = 0; hk < 4; ++hk) {
                                CD_UMMA(d, tc::desc_adv(da, shift + hk * 32), tc::desc_adv(dw, tap * COUT * 128 + hk * 32),
                                              IDESC, accum);
                                accum = 1;
                            }
                        }
                    }
                    CD_COMMIT(&empty_bar[bi]);
                }
                CD_COMMIT(&tfull_bar[acc]);
            }
        }
    } else {
        // ======================= epilogue: + bias, ReLU, bf16 rows in the PS layout of the output ======
        const int lane_grp = warp & 3;
        int it = 0;
        for (int t_ = blockIdx.x; t_ < n_tiles; t_ += gridDim.x, ++it) {
                [[maybe_unused]] const int tile = REV_ORDER ? n_tiles - 1 - t_ : t_;
            const int acc = it & 1;
            const int64_t mp = (int64_t)tile * 128 + lane_grp * 32 + lane;
            const int clip = (int)(mp / P2::RPC), rem = (int)(mp - (int64_t)clip * P2::RPC);
            const int oy = rem / P2::B2, ox = rem - oy * P2::B2;
            const bool live = clip < a.n && oy < HO && ox < WO;
            const int64_t rg = (int64_t)clip * PSO::RPC + (oy + 1) * PSO::WP + (ox + 1);
            tc::mbar_wait(&tfull_bar[acc], (uint32_t)((it >> 1) & 1));
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + acc * COUT;
#pragma unroll 1
            for (int c0 = 0; c0 < COUT; c0 += 32) {
                float v[32];
                tc::tmem_ld32(taddr + c0, v);
                if (c0 + 32 >= COUT) {                 // last read of this accumulator: hand it back early
                    tc::tc_fence_before();
                    tc::mbar_arrive(&tempty_bar[acc]);
                }
                if (live) {
                    __nv_bfloat16* prow = a.y + (int64_t)(c0 / 64) * a.out_plane_elems + rg * 64;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t w[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c = 8 * q + 2 * e;
                            __nv_bfloat162 pk = __floats2bfloat162_rn(fmaxf(v[c] + s_bias[c0 + c], 0.f),
                                                                      fmaxf(v[c + 1] + s_bias[c0 + c + 1], 0.f));
                            w[e] = *reinterpret_cast<uint32_t*>(&pk);
                        }
                        const int jch = (c0 & 63) / 8 + q;
                        *reinterpret_cast<uint4*>(prow + ((jch ^ (int)(rg & 7)) << 3)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ----------------------------------------------------------------------------------------------
// stem on tcgen05: 7x7/s2 conv (1 -> 32) + bias + ReLU + MaxPool2d(2) as ONE GEMM per pooled pixel
// ----------------------------------------------------------------------------------------------
// Row m' = (clip, pooled pixel (py, px)).  The 2x2 pool window's four conv outputs share the 9 x 9 input
// patch at rows 4py-3.., cols 4px-3.., so N = 4 pool members x 32 channels = 128 and the max-pool is an
// in-register max over four 32-column groups of the accumulator row; the un-pooled (32, 45, 51) activation
// never exists.  The patch is covered by 3 x 3 blocks of 4 x 4 pixels, and the feature kernel writes the
// clip as exactly those blocks (cd_common.cuh "blocked map": one 32-byte SWIZZLE_32B row per block, row
// index a * 27 + b), so K = 9 blocks x 16 pixels = 144 and the A operand of block (i, j) is the staged rows
// viewed (27 i + j) rows further down: no A tile is ever built -- one 5.9 KB bulk copy per 128 pooled pixels.
// NF = 90: 22 x 25 = 550 pooled pixels per clip of 648 GEMM rows; NF = 103 (use_delta_delta): 26 x 25 = 650 of 760.
constexpr int kStemRowsIn = (128 + 2 * kBmapB + 2 + 7) & ~7;      // 184 rows a tile reads
constexpr int kStemABytes = ((kStemRowsIn * 32 + 255) / 256) * 256;   // 5888 -> 6144 per stage
constexpr int kStemStages = 4;
constexpr int kStemBBytes = 3 * 16384;            // 128 weight rows x 3 chunks of 64 K-elements (144 used)
constexpr size_t kStemSmem = 1024 + kStemStages * kStemABytes + kStemBBytes + 128;
// warp 0: producer, warp 1: MMA, warps 2-5: epilogue of the even tiles (accumulator 0), warps 6-9: of the odd tiles (accumulator 1).
// One epilogue warp per scheduler spends ~ 760 of a tile's ~ 1 150 cycles in its own dependent instruction stream (ncu stall
// samples, profiles/r2_stem_prune_ab.txt) -- as long as the nine MMAs take; two sets take it off the critical path.
constexpr int kStemThreads = 320;

template <int NF>
__global__ void __launch_bounds__(kStemThreads, 1)
stem_tc_kernel(const __nv_bfloat16* __restrict__ maps, const __nv_bfloat16* __restrict__ wq,
               const float* __restrict__ bias, __nv_bfloat16* __restrict__ x0, int64_t x0_plane_elems,
               __nv_bfloat16* __restrict__ xs0, int B, int rev) {
    const bool REV_ORDER = (rev & 1) != 0;
    const bool FULL_N = (rev & 2) != 0;                         // A/B switch: every tap at N = 128 (CD_STEM_PRUNE=0)
    using NS = NetShape<NF>;
    constexpr int kRows = NS::BROWS;                            // GEMM rows per clip
    extern __shared__ unsigned char raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sA = base;                                   // [kStemStages][kStemABytes]
    unsigned char* sB = base + kStemStages * kStemABytes;       // [3][16384]
    float* s_bias = reinterpret_cast<float*>(sB + kStemBBytes);
    __shared__ uint64_t full_bar[kStemStages], empty_bar[kStemStages], tfull_bar[2], tempty_bar[2], w_bar;
    __shared__ uint32_t tmem_base_s;
    constexpr uint32_t IDESC = tc::idesc_bf16(128, 128), IDESC64 = tc::idesc_bf16(128, 64), IDESC32 = tc::idesc_bf16(128, 32);
    constexpr int TMEM_COLS = 256;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t rows = (int64_t)B * kRows;
    const int n_tiles = (int)((rows + 127) / 128);
    if (tid == 0) {
        for (int i = 0; i < kStemStages; ++i) {
            tc::mbar_init(&full_bar[i], 1);
            tc::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&tfull_bar[i], 1);
            tc::mbar_init(&tempty_bar[i], 128);
        }
        tc::mbar_init(&w_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) s_bias[tid] = bias[tid];
    if (warp == 1) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    tc::pdl_launch_dependents();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ======================= producer: weights once, one bulk copy per tile =====================
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(&w_bar, kStemBBytes);
            tc::bulk_g2s(sB, wq, kStemBBytes, &w_bar);
            tc::pdl_wait();
            int it = 0;
            for (int t_ = blockIdx.x; t_ < n_tiles; t_ += gridDim.x, ++it) {
                [[maybe_unused]] const int tile = REV_ORDER ? n_tiles - 1 - t_ : t_;
                const int s = it % kStemStages;
                tc::mbar_wait(&empty_bar[s], (uint32_t)(((it / kStemStages) & 1) ^ 1));
                tc::mbar_arrive_expect_tx(&full_bar[s], kStemRowsIn * 32);
                tc::bulk_g2s(sA + s * kStemABytes, maps + (int64_t)tile * 128 * 16, kStemRowsIn * 32, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer: nine 4x4-block taps as shifted SWIZZLE_32B views =========
        if (CD_MMA_ISSUER(lane)) {
            tc::mbar_wait(&w_bar, 0);
            const uint64_t db = tc::desc_sw128(tc::s32(sB));
            int it = 0;
            for (int t_ = blockIdx.x; t_ < n_tiles; t_ += gridDim.x, ++it) {
                [[maybe_unused]] const int tile = REV_ORDER ? n_tiles - 1 - t_ : t_;
                const int s = it % kStemStages, acc = it & 1;
                tc::mbar_wait(&tempty_bar[acc], (uint32_t)(((it >> 1) & 1) ^ 1));
                tc::mbar_wait(&full_bar[s], (uint32_t)((it / kStemStages) & 1));
                tc::tc_fence_after();
                const uint32_t d = tmem_base + acc * 128;
                const uint64_t da = tc::desc_sw32(tc::s32(sA + s * kStemABytes));
                // Blocks i = 2 / j = 2 hold only patch row 8 / column 8, which only the pool members with dy = 1 / dx = 1
                // read: their other weight rows are zero.  The members sit in the order (0,0) (0,1) (1,1) (1,0), so those
                // taps run at N = 64 over accumulator columns [64, 128) / [32, 96) and the corner block at N = 32 over
                // [64, 96): 1 952 instead of 2 304 operand rows fetched per tile, same sums (the skipped terms are + 0).
                if (FULL_N) {
#pragma unroll
                    for (int q = 0; q < 9; ++q)                  // block (i, j) = (q / 3, q % 3); K step q of the weights
                        CD_UMMA(d, tc::desc_adv(da, ((q / 3) * kBmapB + (q % 3)) * 32),
                                      tc::desc_adv(db, (q >> 2) * 16384 + (q & 3) * 32), IDESC, q != 0);
                } else {
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const int bi = q / 3, bj = q % 3;
                        const int n0 = (bi == 2) ? 64 : (bj == 2 ? 32 : 0);        // first weight row / accumulator column
                        const uint32_t idesc = (bi == 2 && bj == 2) ? IDESC32 : ((bi == 2 || bj == 2) ? IDESC64 : IDESC);
                        CD_UMMA(d + n0, tc::desc_adv(da, (bi * kBmapB + bj) * 32),
                                      tc::desc_adv(db, (q >> 2) * 16384 + (q & 3) * 32 + n0 * 128), idesc, q != 0);
                    }
                }
                CD_COMMIT(&empty_bar[s]);
                CD_COMMIT(&tfull_bar[acc]);
            }
        }
    } else {
        // ======================= epilogue: pool max over the 4 column groups, bias, ReLU, bf16 ========
        const int lane_grp = warp & 3;                          // TMEM lane quadrant a warp may read = warp % 4
        const int eset = (warp - 2) >> 2;                       // 0: even tiles / accumulator 0, 1: odd tiles / accumulator 1
        int it = 0;
        for (int t_ = blockIdx.x; t_ < n_tiles; t_ += gridDim.x, ++it) {
            if ((it & 1) != eset) continue;
            [[maybe_unused]] const int tile = REV_ORDER ? n_tiles - 1 - t_ : t_;
            const int acc = eset;
            const int mp = tile * 128 + lane_grp * 32 + lane;   // a launch covers <= kTcChunk clips: rows < 2^31
            const int clip = mp / kRows, rem = mp - clip * kRows;
            const int py = rem / kBmapB, px = rem - py * kBmapB;
            const bool live = clip < B && py < NS::H0 && px < NS::W0;
            tc::mbar_wait(&tfull_bar[acc], (uint32_t)((it >> 1) & 1));
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + acc * 128;
            float m[32];
            {   // the four pool members' columns: four loads in flight under one wait, then one 4-way max per channel
                uint32_t r0[32], r1[32], r2[32], r3[32];
                tc::tmem_ld32_issue(taddr, r0);
                tc::tmem_ld32_issue(taddr + 32, r1);
                tc::tmem_ld32_issue(taddr + 64, r2);
                tc::tmem_ld32_issue(taddr + 96, r3);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    m[i] = fmaxf(fmaxf(__uint_as_float(r0[i]), __uint_as_float(r1[i])),
                                 fmaxf(__uint_as_float(r2[i]), __uint_as_float(r3[i])));
            }
            tc::tc_fence_before();
            tc::mbar_arrive(&tempty_bar[acc]);
            if (live) {
                // x0 goes out phase-split (Ps2Layout<22, 25>): plane = (py+1) & 1, row = ((py+1)/2, (px+1)/2),
                // the (px+1) parity picks the 64-byte half of the 128-byte row
                using P2 = Ps2Layout<NS::H0, NS::W0>;
                using PSN = PsLayout<NS::H1, NS::W1>;
                const int64_t r2 = (int64_t)clip * P2::RPC + ((py + 1) >> 1) * P2::B2 + ((px + 1) >> 1);
                __nv_bfloat16* op = x0 + (int64_t)((py + 1) & 1) * x0_plane_elems + r2 * 64;
                const int jb = ((px + 1) & 1) * 4;
                // block 0's skip conv reads x0 at (2y, 2x): emit those pixels as PS rows (32 of 64 channels used)
                const bool xs = !(py & 1) && !(px & 1);
                const int64_t rg = (int64_t)clip * PSN::RPC + (py / 2 + 1) * PSN::WP + (px / 2 + 1);
                __nv_bfloat16* xrow = xs0 + rg * 64;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t w[4];
#pragma unroll
                    const float4 b0 = *reinterpret_cast<const float4*>(s_bias + 8 * q), b1 = *reinterpret_cast<const float4*>(s_bias + 8 * q + 4);
                    const float bq[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};     // folded bias (broadcast reads)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = 8 * q + 2 * e;
                        __nv_bfloat162 pk = __floats2bfloat162_rn(fmaxf(m[c] + bq[2 * e], 0.f), fmaxf(m[c + 1] + bq[2 * e + 1], 0.f));
                        w[e] = *reinterpret_cast<uint32_t*>(&pk);
                    }
                    *reinterpret_cast<uint4*>(op + (((jb + q) ^ (int)(r2 & 7)) << 3)) = make_uint4(w[0], w[1], w[2], w[3]);
                    if (xs) *reinterpret_cast<uint4*>(xrow + ((q ^ (int)(rg & 7)) << 3)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// fp32 features (B, 90, 101) -> the stem's blocked bf16 map (for callers that hand features in): one thread
// per 4x4 block = one 32-byte operand row
template <int NF>
__global__ void __launch_bounds__(256)
map_from_feat_kernel(const float* __restrict__ feat, __nv_bfloat16* __restrict__ maps, int64_t B) {
    constexpr int kRows = NetShape<NF>::BROWS;
    const int64_t b = blockIdx.y;
    const int row = blockIdx.x * 256 + threadIdx.x;
    if (row >= kRows) return;
    const float* f = feat + b * (int64_t)NetShape<NF>::FEAT_ELEMS;
    const int ba = row / kBmapB, bb = row - ba * kBmapB;
    uint32_t w[8];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int fr = 4 * ba + r - 3;
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int t = 4 * bb + c - 3;
            v[c] = (fr >= 0 && fr < NF && t >= 0 && t < kFrames) ? f[fr * kFrames + t] : 0.f;
        }
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
        w[2 * r] = *reinterpret_cast<uint32_t*>(&p0);
        w[2 * r + 1] = *reinterpret_cast<uint32_t*>(&p1);
    }
    uint4* dst = reinterpret_cast<uint4*>(maps + (b * kRows + row) * 16);
    const int sw = (row >> 2) & 1;                   // chunk (r >> 1) lives at (r >> 1) ^ ((row >> 2) & 1)
    dst[0 ^ sw] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1 ^ sw] = make_uint4(w[4], w[5], w[6], w[7]);
}

// AdaptiveAvgPool2d(1) + Flatten + Linear(128 -> 2) from the block-1 kernel's image-row sums psum [B][H][128] fp32
// (NPOS = H * W pixels per clip); fixed summation order, two clips per 256-thread CTA
template <int H, int NPOS>
__global__ void __launch_bounds__(256)
head_kernel(const float* __restrict__ psum, const float* __restrict__ fcw, const float* __restrict__ fcb,
            float* __restrict__ logits, int B) {
    __shared__ float part[2][2][4];
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const int half = threadIdx.x >> 7, c = threadIdx.x & 127;
    const int b = blockIdx.x * 2 + half;
    float a0 = 0.f, a1 = 0.f;
    if (b < B) {
        const float* p = psum + (int64_t)b * H * 128 + c;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < H; ++i) s += p[i * 128];
        s *= (1.0f / (float)NPOS);
        a0 = s * fcw[c];
        a1 = s * fcw[128 + c];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    if ((c & 31) == 0) {
        part[half][0][c >> 5] = a0;
        part[half][1][c >> 5] = a1;
    }
    __syncthreads();
    if (c < 2 && b < B)
        logits[(int64_t)b * 2 + c] = ((part[half][c][0] + part[half][c][1]) + (part[half][c][2] + part[half][c][3])) + fcb[c];
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
// folded weights [tap][cin][cout] fp32 -> [kc][cout rows x 128 B] bf16, SWIZZLE_128B K-major, K padded to 64
static void pack_weights(const std::vector<float>& wf, int taps, int cin, int cout, std::vector<__nv_bfloat16>& out) {
    const int ktot = taps * cin, nkc = (ktot + 63) / 64;
    out.assign((size_t)nkc * cout * 64, __float2bfloat16_rn(0.f));
    for (int n = 0; n < cout; ++n)
        for (int k = 0; k < ktot; ++k) {
            const int kc = k / 64, kk = k % 64, ch = kk / 8, e = kk % 8;
            const size_t off = (size_t)kc * cout * 64 + (size_t)n * 64 + (size_t)((ch ^ (n & 7)) * 8) + e;
            out[off] = __float2bfloat16_rn(wf[(size_t)k * cout + n]);      // k = tap*cin + ci
        }
}

template <int NF> static int init_cnn_tc_nf(cd_handle* h);
int init_cnn_tc(cd_handle* h) {
    CD_CUDA(h, cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CD_CUDA(h, cudaFuncSetAttribute(umma_pair_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int rc = init_cnn_tc_nf<kFeat>(h);
    return rc != CD_OK ? rc : init_cnn_tc_nf<kFeatDD>(h);
}

// wf/bf: [stem, b0c1, b0sk, b0c2, b1c1, b1sk, b1c2, fc] folded fp32, conv weights as [tap][cin][cout]
int upload_cnn_tc(cd_handle* h, const std::vector<std::vector<float>>& wf, const std::vector<std::vector<float>>& bf) {
    std::vector<__nv_bfloat16> packs[7];
    const int spec[6][3] = {{9, 32, 64}, {1, 32, 64}, {9, 64, 64}, {9, 64, 128}, {1, 64, 128}, {9, 128, 128}};
    size_t total_bf = 0;
    for (int i = 0; i < 6; ++i) {
        pack_weights(wf[1 + i], spec[i][0], spec[i][1], spec[i][2], packs[i]);
        total_bf += (packs[i].size() + 63) & ~size_t(63);
    }
    {   // stem as a pooled GEMM over 3x3 blocks of 4x4 pixels: patch pixel (ry, rx) = (2dy + ky, 2dx + kx) lives in
        // block (ry / 4, rx / 4) at k16 = 4 (ry % 4) + rx % 4:  W2[k][n = member(dy,dx)*32 + co] = w[co][ky][kx], members in the
        // order (0,0) (0,1) (1,1) (1,0) so that "dy = 1" and "dx = 1" are each a contiguous range of weight rows (stem kernel)
        std::vector<float> w2((size_t)192 * 128, 0.f);
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx)
                for (int ky = 0; ky < 7; ++ky)
                    for (int kx = 0; kx < 7; ++kx)
                        for (int co = 0; co < 32; ++co) {
                            const int ry = 2 * dy + ky, rx = 2 * dx + kx;
                            const int k = ((ry >> 2) * 3 + (rx >> 2)) * 16 + (ry & 3) * 4 + (rx & 3);
                            const int member = dy ? 3 - dx : dx;
                            w2[(size_t)k * 128 + member * 32 + co] = wf[0][(size_t)(ky * 7 + kx) * 32 + co];
                        }
        pack_weights(w2, 192, 1, 128, packs[6]);
        total_bf += (packs[6].size() + 63) & ~size_t(63);
    }
    std::vector<float> fblob;
    auto push = [&](const std::vector<float>& v) {
        size_t off = fblob.size();
        fblob.insert(fblob.end(), v.begin(), v.end());
        while (fblob.size() % 4) fblob.push_back(0.f);
        return off;
    };
    // the skip projection is accumulated into conv2's tile (tenth tap), so its bias folds into conv2's
    std::vector<float> b0c2s(bf[3]), b1c2s(bf[6]);
    for (size_t i = 0; i < b0c2s.size(); ++i) b0c2s[i] += bf[2][i];
    for (size_t i = 0; i < b1c2s.size(); ++i) b1c2s[i] += bf[5][i];
    const size_t o_stem_b = push(bf[0]), o_b0c1 = push(bf[1]), o_b0sk = push(bf[2]),
                 o_b0c2 = push(b0c2s), o_b1c1 = push(bf[4]), o_b1sk = push(bf[5]), o_b1c2 = push(b1c2s),
                 o_fcw = push(wf[7]), o_fcb = push(bf[7]);
    const size_t bf_bytes = total_bf * sizeof(__nv_bfloat16), f_bytes = fblob.size() * sizeof(float);
    if (!h->d_wblob_bf16) CD_CUDA(h, cudaMalloc(&h->d_wblob_bf16, bf_bytes + f_bytes));
    std::vector<unsigned char> host(bf_bytes + f_bytes, 0);
    __nv_bfloat16* dptr[7];
    size_t off = 0;
    for (int i = 0; i < 7; ++i) {
        memcpy(host.data() + off * 2, packs[i].data(), packs[i].size() * 2);
        dptr[i] = h->d_wblob_bf16 + off;
        off += (packs[i].size() + 63) & ~size_t(63);
    }
    memcpy(host.data() + bf_bytes, fblob.data(), f_bytes);
    CD_CUDA(h, cudaMemcpy(h->d_wblob_bf16, host.data(), host.size(), cudaMemcpyHostToDevice));
    float* fbase = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(h->d_wblob_bf16) + bf_bytes);
    TcWeights& W = h->tcw;
    W.b0c1 = dptr[0]; W.b0sk = dptr[1]; W.b0c2 = dptr[2]; W.b1c1 = dptr[3]; W.b1sk = dptr[4]; W.b1c2 = dptr[5];
    W.stem_w = dptr[6]; W.stem_b = fbase + o_stem_b; W.b0c1_b = fbase + o_b0c1; W.b0sk_b = fbase + o_b0sk;
    W.b0c2_b = fbase + o_b0c2; W.b1c1_b = fbase + o_b1c1; W.b1sk_b = fbase + o_b1sk; W.b1c2_b = fbase + o_b1c2;
    W.fc_w = fbase + o_fcw; W.fc_b = fbase + o_fcb;
    return CD_OK;
}

// Tensor-core workspace: fixed layout for kTcChunk clips of one shape family, allocated once per handle for the larger
// family and zeroed whenever the family changes (the PS buffers rely on their pad rows staying zero, and the two
// families put their pads in different places).
// [maps | x0 | h1 PS: conv plane + skip plane | x1 | h2 PS: 2 conv planes + skip plane | row sums of x2]
constexpr int64_t kTcChunk = 16384;
constexpr int64_t al256(int64_t x) { return (x + 255) & ~int64_t(255); }
template <int NF>
struct TcPlan {
    using N = NetShape<NF>;
    using Ps0 = PsLayout<N::H1, N::W1>;
    using Ps1 = PsLayout<N::H2, N::W2>;
    using P20 = Ps2Layout<N::H0, N::W0>;
    using P21 = Ps2Layout<N::H1, N::W1>;
    static constexpr int NPOS = N::H2 * N::W2;
    static constexpr int64_t kPs0Rows = ((kTcChunk * Ps0::RPC + 127) / 128) * 128 + Ps0::ROWS_IN;          // + the last tile's look-ahead
    static constexpr int64_t kPs1Rows = ((kTcChunk * Ps1::RPC + 127) / 128) * 128 + 128 + Ps1::ROWS_IN;    // + the pair kernel's odd tile
    static constexpr int64_t kP20Rows = ((kTcChunk * P20::RPC + 127) / 128) * 128 + P20::ROWS_IN;          // x0, 2 planes
    static constexpr int64_t kP21Rows = ((kTcChunk * P21::RPC + 127) / 128) * 128 + P21::ROWS_IN;          // x1, 4 planes
    static constexpr int64_t kOffMap = 0;
    static constexpr int64_t kOffX0 = kOffMap + al256((kTcChunk * N::BROWS + 128 + kStemRowsIn) * 32);     // + last tile's look-ahead
    static constexpr int64_t kOffH1 = kOffX0 + al256(2 * kP20Rows * 128);
    static constexpr int64_t kOffX1 = kOffH1 + al256(2 * kPs0Rows * 128);
    static constexpr int64_t kOffH2 = kOffX1 + al256(4 * kP21Rows * 128);
    static constexpr int64_t kOffX2 = kOffH2 + al256(3 * kPs1Rows * 128);
    static constexpr int64_t kBytes = kOffX2 + al256(kTcChunk * N::H2 * 128 * 4);      // image-row sums of the block-1 output
};
constexpr int64_t kTcWsBytes = TcPlan<kFeatDD>::kBytes > TcPlan<kFeat>::kBytes ? TcPlan<kFeatDD>::kBytes : TcPlan<kFeat>::kBytes;

size_t cnn_tc_ws_bytes(int64_t) { return 0; }      // nothing from the shared workspace

// nf: the shape family (90 | 103 feature rows) the next launches will use.  Each family has its own workspace
// (allocated and zeroed on first use, 2.9 / 3.4 GB of 180), so switching families costs nothing and calls of
// different families never share padding.
int ensure_tc_ws(cd_handle* h, int nf) {
    const int fam = nf == kFeatDD ? 1 : 0;
    if (!h->d_ws_tc_fam[fam]) {
        const size_t bytes = (size_t)(fam ? TcPlan<kFeatDD>::kBytes : TcPlan<kFeat>::kBytes);
        cudaError_t e = cudaMalloc(&h->d_ws_tc_fam[fam], bytes);
        if (e != cudaSuccess) return set_err(h, CD_E_NOMEM, "tensor-core workspace allocation", e);
        CD_CUDA(h, cudaMemset(h->d_ws_tc_fam[fam], 0, bytes));
        CD_CUDA(h, cudaDeviceSynchronize());       // callers' streams are non-blocking: order the zero-fill before them
    }
    h->d_ws_tc = h->d_ws_tc_fam[fam];
    return CD_OK;
}

// the first region holds the bf16 maps of the current chunk: the feature kernel writes them there
// directly in the end-to-end path (call ensure_tc_ws first)
__nv_bfloat16* cnn_tc_map_ptr(cd_handle* h) { return reinterpret_cast<__nv_bfloat16*>(h->d_ws_tc); }

template <int NF>
static int init_cnn_tc_nf(cd_handle* h) {
    using N = NetShape<NF>;
    CD_CUDA(h, cudaFuncSetAttribute(conv3x3_s2_tc_kernel<32, 64, N::H0, N::W0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)S2Shape<32, 64, N::H0, N::W0>::SMEM));
    CD_CUDA(h, cudaFuncSetAttribute(conv3x3_s2_tc_kernel<64, 128, N::H1, N::W1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)S2Shape<64, 128, N::H1, N::W1>::SMEM));
    CD_CUDA(h, cudaFuncSetAttribute(conv3x3_s1_tc_kernel<64, N::H1, N::W1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)S1Shape<64, N::H1, N::W1>::SMEM));
    CD_CUDA(h, cudaFuncSetAttribute(conv3x3_s1_pair_tc_kernel<N::H2, N::W2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)S1PairShape<N::H2, N::W2>::SMEM));
    CD_CUDA(h, cudaFuncSetAttribute(stem_tc_kernel<NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStemSmem));
    return CD_OK;
}

// Launch with programmatic stream serialization (see tc::pdl_wait): the kernel may be scheduled while its predecessor
// in the stream drains; its own griddepcontrol.wait orders the data.
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// d_feat != nullptr: features come in as fp32 and are converted chunk by chunk; d_feat == nullptr:
// the maps of the (single, <= kTcChunk clips) chunk are already in the workspace.
template <int NF>
static int launch_cnn_tc_nf(cd_handle* h, const float* d_feat, int64_t B, float* d_logits, cudaStream_t s) {
    using N = NetShape<NF>;
    using P = TcPlan<NF>;
    const TcWeights& W = h->tcw;
    int rc0 = ensure_tc_ws(h, NF);
    if (rc0 != CD_OK) return rc0;
    unsigned char* ws = h->d_ws_tc;
    __nv_bfloat16* maps = reinterpret_cast<__nv_bfloat16*>(ws + P::kOffMap);
    __nv_bfloat16* x0 = reinterpret_cast<__nv_bfloat16*>(ws + P::kOffX0);
    __nv_bfloat16* h1 = reinterpret_cast<__nv_bfloat16*>(ws + P::kOffH1);   // plane 0: conv1 output, plane 1: x0 at (2y, 2x)
    __nv_bfloat16* x1 = reinterpret_cast<__nv_bfloat16*>(ws + P::kOffX1);
    __nv_bfloat16* h2 = reinterpret_cast<__nv_bfloat16*>(ws + P::kOffH2);   // planes 0, 1: conv1 output, plane 2: x1 at (2y, 2x)
    float* x2 = reinterpret_cast<float*>(ws + P::kOffX2);
    constexpr int64_t kPl0 = P::kPs0Rows * 64, kPl1 = P::kPs1Rows * 64;
    // Tile order alternates down the chain: a kernel that walks its tiles backwards starts with the rows its producer
    // wrote LAST, which are still dirty in L2 (the hand-offs of a 16 384-clip slab are 3 - 5 x the L2, so only that end
    // can hit).  feature kernel / map conversion: forward;  stem: backward;  b0 conv1: forward;  b0 conv2: backward;
    // b1 conv1: forward;  b1 conv2: backward.  CD_TC_ORDER=0 walks every kernel forward (the A/B switch of tools/slab_sweep.py).
    const char* order_env = getenv("CD_TC_ORDER");
    const int alt = (order_env && order_env[0] == '0') ? 0 : 1;
    const char* prune_env = getenv("CD_STEM_PRUNE");            // CD_STEM_PRUNE=0: every stem tap at N = 128 (A/B switch)
    const int stem_full = (prune_env && prune_env[0] == '0') ? 2 : 0;
    for (int64_t c0 = 0; c0 < B; c0 += kTcChunk) {
        const int n = (int)((B - c0) < kTcChunk ? (B - c0) : kTcChunk);
        if (d_feat) {
            map_from_feat_kernel<NF><<<dim3((N::BROWS + 255) / 256, n), 256, 0, s>>>(d_feat + c0 * N::FEAT_ELEMS, maps, n);
            h->launches += 1;
        }
        {
            const int tiles = (int)(((int64_t)n * N::BROWS + 127) / 128);
            CD_CUDA(h, launch_pdl(stem_tc_kernel<NF>, dim3(tiles < h->sm_count ? tiles : h->sm_count), dim3(kStemThreads), kStemSmem, s,
                                  (const __nv_bfloat16*)maps, (const __nv_bfloat16*)W.stem_w, (const float*)W.stem_b, x0,
                                  (int64_t)(P::kP20Rows * 64), h1 + kPl0, n, alt | stem_full));
        }
        h->launches += 1;
        {
            using S = S2Shape<32, 64, N::H0, N::W0>;
            S2Args q{};
            q.x = x0; q.w = W.b0c1; q.bias = W.b0c1_b; q.y = h1; q.in_plane_elems = P::kP20Rows * 64; q.out_plane_elems = kPl0; q.n = n; q.rev = 0;
            const int tiles = (int)(((int64_t)n * P::P20::RPC + 127) / 128);
            CD_CUDA(h, launch_pdl(conv3x3_s2_tc_kernel<32, 64, N::H0, N::W0>, dim3(tiles < h->sm_count ? tiles : h->sm_count),
                                  dim3(kS1Threads), S::SMEM, s, q));
            h->launches += 1;
        }
        {
            using S = S1Shape<64, N::H1, N::W1>;
            S1Args q{};
            q.x = h1; q.w = W.b0c2; q.wskip = W.b0sk; q.bias = W.b0c2_b; q.y_bf16 = x1; q.xs_out = h2 + 2 * kPl1;
            q.y_plane_elems = P::kP21Rows * 64; q.ps_plane_elems = kPl0; q.n = n; q.rev = alt;
            const int tiles = (int)(((int64_t)n * P::Ps0::RPC + 127) / 128);
            const int grid = (tiles < h->sm_count ? tiles : h->sm_count) * S::HALVES;
            CD_CUDA(h, launch_pdl(conv3x3_s1_tc_kernel<64, N::H1, N::W1, false, true>, dim3(grid), dim3(kS1Threads), S::SMEM, s, q));
            h->launches += 1;
        }
        {
            using S = S2Shape<64, 128, N::H1, N::W1>;
            S2Args q{};
            q.x = x1; q.w = W.b1c1; q.bias = W.b1c1_b; q.y = h2; q.in_plane_elems = P::kP21Rows * 64; q.out_plane_elems = kPl1; q.n = n; q.rev = 0;
            const int tiles = (int)(((int64_t)n * P::P21::RPC + 127) / 128);
            CD_CUDA(h, launch_pdl(conv3x3_s2_tc_kernel<64, 128, N::H1, N::W1>, dim3(tiles < h->sm_count ? tiles : h->sm_count),
                                  dim3(kS1Threads), S::SMEM, s, q));
            h->launches += 1;
        }
        {
            using S = S1PairShape<N::H2, N::W2>;
            S1Args q{};
            q.x = h2; q.w = W.b1c2; q.wskip = W.b1sk; q.bias = W.b1c2_b; q.psum = x2; q.ps_plane_elems = kPl1; q.n = n; q.rev = alt;
            const int tiles = (int)(((int64_t)n * P::Ps1::RPC + 127) / 128);
            const int tp = (tiles + 1) / 2, max_pairs = h->sm_count / 2;
            CD_CUDA(h, launch_pdl(conv3x3_s1_pair_tc_kernel<N::H2, N::W2>, dim3(2 * (tp < max_pairs ? tp : max_pairs)),
                                  dim3(kS1Threads), S::SMEM, s, q));
            h->launches += 1;
        }
        CD_CUDA(h, launch_pdl(head_kernel<N::H2, P::NPOS>, dim3((n + 1) / 2), dim3(256), 0, s, (const float*)x2, (const float*)W.fc_w,
                              (const float*)W.fc_b, d_logits + c0 * CD_N_CLASSES, n));
        h->launches += 1;
        CD_CUDA(h, cudaGetLastError());
    }
    return CD_OK;
}

int launch_cnn_tc(cd_handle* h, const float* d_feat, int64_t B, float* d_logits, cudaStream_t s, int nf) {
    if (B == 0) return CD_OK;
    if (!d_feat && B > kTcChunk) return set_err(h, CD_E_ARG, "launch_cnn_tc: pre-filled maps hold one chunk only");
    return nf == kFeatDD ? launch_cnn_tc_nf<kFeatDD>(h, d_feat, B, d_logits, s) : launch_cnn_tc_nf<kFeat>(h, d_feat, B, d_logits, s);
}

// D = A * B^T for 128 x K by N x K bf16-rounded inputs (host fp32 in/out); exercises descriptor encodings
int umma_selftest(cd_handle* h, int N, int K, int shift, int mode, const float* hA, const float* hB, float* hD) {
    if (mode == 3) {        // CTA pair: A is 256 x K, D is 256 x N
        const size_t smem2 = (size_t)(128 + N / 2) * K * 2 + 2048;
        if (N < 32 || N > 256 || N % 32 || K < 64 || K % 64 || shift != 0 || smem2 > 200 * 1024)
            return set_err(h, CD_E_ARG, "umma_selftest(pair): need N in [32,256] step 32, K multiple of 64, shift 0");
        std::vector<__nv_bfloat16> a((size_t)256 * K), b((size_t)N * K);
        for (size_t i = 0; i < a.size(); ++i) a[i] = __float2bfloat16_rn(hA[i]);
        for (size_t i = 0; i < b.size(); ++i) b[i] = __float2bfloat16_rn(hB[i]);
        __nv_bfloat16 *dA = nullptr, *dB = nullptr;
        float* dD = nullptr;
        CD_CUDA(h, cudaMalloc(&dA, a.size() * 2));
        CD_CUDA(h, cudaMalloc(&dB, b.size() * 2));
        CD_CUDA(h, cudaMalloc(&dD, (size_t)256 * N * 4));
        CD_CUDA(h, cudaMemcpy(dA, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
        CD_CUDA(h, cudaMemcpy(dB, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
        umma_pair_selftest_kernel<<<2, 128, smem2>>>(dA, dB, dD, N, K);
        h->launches += 1;
        CD_CUDA(h, cudaGetLastError());
        CD_CUDA(h, cudaDeviceSynchronize());
        CD_CUDA(h, cudaMemcpy(hD, dD, (size_t)256 * N * 4, cudaMemcpyDeviceToHost));
        cudaFree(dA);
        cudaFree(dB);
        cudaFree(dD);
        return CD_OK;
    }
    const size_t smem = (size_t)(136 + shift + N) * K * 2 + 2048;
    if (N < 32 || N > 256 || N % 32 || K < 64 || K % 64 || shift < 0 || shift > 64 || smem > 200 * 1024)
        return set_err(h, CD_E_ARG, "umma_selftest: need N in [32,256] step 32, K multiple of 64, shift in [0,64], tiles within 200 KB");
    std::vector<__nv_bfloat16> a((size_t)(128 + shift) * K), b((size_t)N * K);
    for (size_t i = 0; i < a.size(); ++i) a[i] = __float2bfloat16_rn(hA[i]);
    for (size_t i = 0; i < b.size(); ++i) b[i] = __float2bfloat16_rn(hB[i]);
    __nv_bfloat16 *dA = nullptr, *dB = nullptr;
    float* dD = nullptr;
    CD_CUDA(h, cudaMalloc(&dA, a.size() * 2));
    CD_CUDA(h, cudaMalloc(&dB, b.size() * 2));
    CD_CUDA(h, cudaMalloc(&dD, (size_t)128 * N * 4));
    CD_CUDA(h, cudaMemcpy(dA, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
    CD_CUDA(h, cudaMemcpy(dB, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
    umma_selftest_kernel<<<1, 128, smem>>>(dA, dB, dD, N, K, shift, mode);
    h->launches += 1;
    CD_CUDA(h, cudaGetLastError());
    CD_CUDA(h, cudaDeviceSynchronize());
    CD_CUDA(h, cudaMemcpy(hD, dD, (size_t)128 * N * 4, cudaMemcpyDeviceToHost));
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dD);
    return CD_OK;
}

}  // namespace cd
