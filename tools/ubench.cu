// Pipe microbenchmarks behind the feature kernel's design: warp-instructions per clock per SM for the
// instruction kinds the FFT is made of, at the kernel's occupancy (16 warps/SM) and at 32 / 48 warps/SM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/ubench.cu -o /tmp/ubench && /tmp/ubench
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long u64;
#define ITER 2048

__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

// mode 0: FFMA (3 reg)  1: FFMA2  2: FADD2  3: FMUL2  4: FADD  5: SHFL  6: LDS.64  7: STS.64+LDS.64  8: LDS.64 + SHFL mixed
// 9: FFMA2 + LDS.64 mixed (1:1)  10: LDS.32  11: LDS.128  12: FMNMX  13: FFMA2 + FADD2 mixed  14: MUFU.LG2
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    __shared__ __align__(16) float sm[256 * 16];        // 16 KB: 8 rows of 256 float2
    const int tid = threadIdx.x;
    for (int i = tid; i < 256 * 16; i += 256) sm[i] = (float)i * 1e-3f;
    __syncthreads();
    float a[8];
    u64 p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = tid * 0.001f + i; p[i] = pk(a[i], a[i] + 0.5f); }
    const float c0 = 1.0001f, c1 = 0.0001f;
    const u64 pc0 = pk(c0, c0), pc1 = pk(c1, c1);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    uint32_t addr = sbase + tid * 8;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], c0, c1);
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc0), "l"(pc1));
            if (MODE == 2) asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc1));
            if (MODE == 3) asm volatile("mul.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc0));
            if (MODE == 4) a[i] = a[i] + c1;
            if (MODE == 5) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 3));
            if (MODE == 6) { float x, y; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(addr + ((it + i) & 7) * 2048)); a[i] += x + y; }
            if (MODE == 7) {
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr + (i & 7) * 2048), "f"(a[i]), "f"(a[i]));
                float x, y; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(addr + ((i + 1) & 7) * 2048)); a[i] = x;
            }
            if (MODE == 8) {
                float x, y; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(addr + ((it + i) & 7) * 2048));
                a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 3)) + x;
            }
            if (MODE == 9) {
                float x, y; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(addr + ((it + i) & 7) * 2048));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc0), "l"(pk(x, y)));
            }
            if (MODE == 10) { float x; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(sbase + tid * 4 + ((it + i) & 7) * 1024)); a[i] += x; }
            if (MODE == 11) { float x, y, z, w; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(sbase + tid * 16 + ((it + i) & 1) * 4096)); a[i] += x + w; }
            if (MODE == 12) a[i] = fmaxf(a[i], c0 + (float)it);
            if (MODE == 13) {
                if (i & 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc0), "l"(pc1));
                else asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc1));
            }
            if (MODE == 14) a[i] = __log2f(a[i]) + 3.f;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + lo(p[i]);
    out[blockIdx.x * 256 + tid] = s;
}

template <int MODE>
static void run(const char* name, int extra_per_iter, float* d_out) {
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    for (int ctas_per_sm : {2, 4, 6}) {
        const int grid = sms * ctas_per_sm;
        k<MODE><<<grid, 256>>>(d_out, 64);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<MODE><<<grid, 256>>>(d_out, ITER);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(err)); fflush(stdout); exit(1); }
        const double winstr = (double)grid * 8 /*warps*/ * ITER * 8 * (1 + extra_per_iter);
        const double clk = ms * 1e-3 * (double)khz * 1e3;
        fflush(stdout);
        printf("%-28s warps/SM %2d  %.3f ms  %.2f warp-instr/clk/SM (counted kind%s)\n", name, ctas_per_sm * 8, ms, winstr / clk / sms,
               extra_per_iter ? ", both kinds" : "");
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
}

int main() {
    float* d_out;
    cudaMalloc(&d_out, 148 * 8 * 256 * 4);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("nominal clock %d kHz (rates below assume it; compare RATIOS)\n", khz);
    run<0>("FFMA 3-reg", 0, d_out);
    run<4>("FADD", 0, d_out);
    run<1>("FFMA2 (fma.f32x2)", 0, d_out);
    run<2>("FADD2 (add.f32x2)", 0, d_out);
    run<3>("FMUL2 (mul.f32x2)", 0, d_out);
    run<13>("FFMA2+FADD2 alternating", 0, d_out);
    run<12>("FMNMX", 0, d_out);
    run<14>("MUFU.LG2 (+FADD)", 1, d_out);
    run<5>("SHFL.BFLY", 0, d_out);
    run<10>("LDS.32 (+FADD)", 1, d_out);
    run<6>("LDS.64 (+2 FADD)", 0, d_out);
    run<11>("LDS.128 (+2 FADD)", 0, d_out);
    run<7>("STS.64 + LDS.64", 1, d_out);
    run<8>("LDS.64 + SHFL", 1, d_out);
    run<9>("LDS.64 + FFMA2", 1, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
