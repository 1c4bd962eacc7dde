// tcgen05.mma issue-rate microbenchmark behind DESIGN.md section 4.2 / 7: what one K = 16, M = 128 bf16 MMA costs as a function of
// N, of the A operand's layout (SWIZZLE_128B rows as in the 3x3 convolutions, SWIZZLE_32B rows as in the stem) and of whether
// consecutive MMAs accumulate into the SAME TMEM accumulator (a dependent chain, as in every conv kernel here) or alternate
// between 2 / 4 accumulators (independent chains).  One CTA per SM, one issuing thread, REPS MMAs then one commit; cycles are
// clock64() around issue + completion, per MMA, median over the CTAs.  Operand contents are zeros: timing is data-independent.
// RESULT (profiles/r2_umma_issue_ubench.txt): 143.9 cycles per MMA in EVERY configuration, issue-only 143.8 -- the loop below (a runtime
// modulo, descriptor arithmetic, the compiler's R2UR / ELECT / BRA.U.ANY wrapper around a tcgen05.mma issued from one thread of a
// divergent region) is slower than the tensor core at any N <= 256, i.e. the benchmark measures the ISSUE path.  That is the finding it
// was kept for; a version that isolates the tensor core needs a convergent, fully unrolled issue sequence.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/umma_bench.cu -o /tmp/umma_bench && /tmp/umma_bench
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {     // rows of 128 B, 8-row groups 1024 B apart
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)64 << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint64_t desc_sw32(uint32_t saddr) {      // rows of 32 B, 8-row groups 256 B apart
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)16 << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// a_sw32: A as [K16 step][136 rows][32 B] (the stem's operand form); otherwise [136 rows][128 B] with the K16 step as a 32-byte slice
__global__ void __launch_bounds__(128) umma_rate(int N, int reps, int n_acc, int a_sw32, int shift_rows, long long* out) {
    extern __shared__ unsigned char raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sA = base;                    // 32 KB: 256 rows x 128 B (room for shifted views)
    unsigned char* sB = base + 32768;            // 32 KB: up to 256 rows x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 16; i += 128) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = idesc_bf16(128, N);
        const uint64_t da = a_sw32 ? desc_sw32(s32(sA)) : desc_sw128(s32(sA));
        const uint64_t db = desc_sw128(s32(sB));
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const int k4 = r & 3, tap = (r >> 2) % 9;                                   // K16 slice, "filter tap" (a shifted view)
            const uint32_t a_off = a_sw32 ? (uint32_t)(k4 * 8192 + tap * shift_rows * 32)
                                          : (uint32_t)(k4 * 32 + tap * shift_rows * 128);
            umma_bf16(tm + (uint32_t)((r % n_acc) * 128), da + (a_off >> 4), db + (uint64_t)((k4 * 32) >> 4), idesc, r >= n_acc);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
        const long long t1 = clock64();
        while (!mbar_try_wait(&bar, 0)) {
            if (clock64() - t1 > 2000000000LL) __trap();
        }
        const long long t2 = clock64();
        out[2 * blockIdx.x] = t2 - t0;
        out[2 * blockIdx.x + 1] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

int main() {
    int dev = 0, sms = 0;
    cudaSetDevice(dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int smem = 65536 + 1024;
    cudaFuncSetAttribute(umma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* d_out;
    cudaMalloc(&d_out, sizeof(long long) * 2 * sms);
    std::vector<long long> h(2 * sms);
    const int reps = 1152;                       // 32 "tiles" of 36 MMAs
    printf("M = 128, K = 16 bf16 tcgen05.mma, %d per CTA, %d CTAs (one per SM); cycles per MMA = median over CTAs of (issue + completion) / count\n", reps, sms);
    printf("%-8s %-6s %-5s %-6s %10s %12s %10s\n", "A", "N", "acc", "shift", "cyc/MMA", "issue-only", "math floor");
    for (int a_sw32 = 0; a_sw32 < 2; ++a_sw32)
        for (int N : {32, 64, 128, 256})
            for (int n_acc : {1, 2, 4})
                for (int shift : {0, 1}) {
                    if (N * n_acc > 512 || (N == 256 && n_acc > 2)) continue;
                    if (n_acc * 128 > 512) continue;
                    if (N > 128 && n_acc > 1) continue;             // accumulators are spaced 128 columns apart
                    for (int w = 0; w < 2; ++w)                     // second launch is the measured one
                        umma_rate<<<sms, 128, smem>>>(N, reps, n_acc, a_sw32, shift, d_out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(h.data(), d_out, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost);
                    std::vector<long long> tot, iss;
                    for (int i = 0; i < sms; ++i) { tot.push_back(h[2 * i]); iss.push_back(h[2 * i + 1]); }
                    std::sort(tot.begin(), tot.end());
                    std::sort(iss.begin(), iss.end());
                    printf("%-8s %-6d %-5d %-6d %10.1f %12.1f %10.1f\n", a_sw32 ? "SW32" : "SW128", N, n_acc, shift,
                           (double)tot[sms / 2] / reps, (double)iss[sms / 2] / reps, 128.0 * N * 16 / 4096.0);
                }
    printf("no error\n");
    return 0;
}
