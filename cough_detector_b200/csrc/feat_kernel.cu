// Fused feature kernel: clip -> (90, 101) features, persistent, two 256-thread CTAs per SM.
//
// Replaces, per clip, AudioPreprocessor.normalize (src/preprocessing.py:199-212), pad_or_trim
// (:358-385) and extract_features (:432-489): T.MelSpectrogram (reflect pad, 400-sample periodic
// Hann centred in a 512-point frame, rFFT, power, 64-band HTK mel), AmplitudeToDB(top_db=80),
// (dB+80)/80 clamp, T.MFCC's DCT-II of the floored dB, global mean / unbiased-std normalisation,
// and the replicate-padded central-difference deltas.  The reference computes the mel
// spectrogram twice (extract_mel_spectrogram and again inside T.MFCC); here it is computed once.
//
// Data flow per clip (all on chip between the 64 000 B read and the feature / operand-map write):
//   global --one cp.async.bulk (TMA 1-D), issued under the previous clip's tail--> smem clip[16000]
//   block max|x| -> 1/max folded into the lane's 13 window weights (the clip is never rewritten)
//   8 warps x frame PAIRS: z = w*(x_2p + i x_2p+1), complex 512-point FFT as 8x8x8 radix-8
//       (16 points per lane in registers, packed f32x2 math, two smem exchanges), Z[k], Z[512-k] -> the two
//       power spectra for bins 4..127 only, segment-owned banded mel, 10 log10 -> smem dB tile [64][101]
//   block max(dB) -> floor; mel01 rows written; DCT-13 from the floored tile; block mean / std;
//   normalised MFCC + deltas written (fp32 features and/or the CNN stem's blocked bf16 map).
// The power spectrum is invariant to a circular shift of the frame, so the 400 windowed samples
// sit at n = 0..399 of the 512-point frame instead of torch.stft's 56..455.
#include "cd_common.cuh"
#include "cd_tables.h"
// Ablation switches (never set in the shipped library; `make ablations` builds build/exp/libexp<n>.so and
// tools/feat_time.py times them -- results are WRONG by construction, only the time is of interest):
//   1 no second shared-memory exchange   2 no exchanges   3 no STFT at all   4 no tail   5 one mel slot per segment
#ifndef CD_EXP
#define CD_EXP 0
#endif

namespace cd {

constexpr int kFeatWarps   = 8;                // two CTAs per SM: one CTA's tail overlaps the other's FFTs
constexpr int kFeatThreads = kFeatWarps * 32;
constexpr int kXbElems     = 288;             // float2 per warp: 4 rows x 72 (64 + 8 pad)
constexpr int kXbRow       = 72;

__constant__ float c_dct[kMfcc * kMels];      // [13][64], ortho DCT-II rows

struct FeatSmem {
    float  clip[kSeg];                        //  64 000 B, filled by one bulk copy per clip
    float2 xb[kFeatWarps][kXbElems];          //  18 432 B, per-warp FFT exchange; MFCC scratch in the tail
    float  db[kMels * kFrames];               //  25 856 B, dB tile [band][frame]
    float2 tw1[7][64];                        //   3 584 B
    float2 tw2[7][8];                         //     448 B
    float2 zero2;                             // what the mel contraction reads for a lane's unused bin slots
    float  red[64];
    unsigned long long bar;
};                                            // 112 592 B -> two CTAs per SM
static_assert(sizeof(FeatSmem) <= 113 * 1024, "two CTAs per SM");
static_assert(offsetof(FeatSmem, xb) % 128 == 0 && offsetof(FeatSmem, zero2) % 128 == 64, "zero2 sits in bank pair 8 of xb's numbering");

// ----------------------------------------------------------------------------------------------
// small PTX wrappers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded like the tensor-core kernels' waits: a bulk copy that never lands traps (-> a launch error the host
// reports) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();     // ~2 s at 1.9 GHz
    }
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}
// 1-D bulk async copy global -> shared (TMA engine), completion signalled on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Packed fp32 pairs (add/sub/mul/fma.f32x2 -> FADD2 / FMUL2 / FFMA2 on sm_100): a complex value is one 64-bit register
// pair, so a complex add is ONE instruction and a complex multiply TWO (ptxas folds the half swaps, the per-half
// negation and the scalar broadcasts into operand modifiers).  Butterflies, twiddles, the window and the power
// spectrum are ~2/3 of the kernel's instructions, which is issue-bound -- this halves them.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 up(u64 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    u64 rd;
    asm("add.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(pk(a.x, a.y)), "l"(pk(b.x, b.y)));
    return up(rd);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    u64 rd;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(pk(a.x, a.y)), "l"(pk(b.x, b.y)));
    return up(rd);
}
// s * a
__device__ __forceinline__ float2 scale2(float s, float2 a) {
    u64 rd;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(pk(s, s)), "l"(pk(a.x, a.y)));
    return up(rd);
}
// s * a + c
__device__ __forceinline__ float2 fma2s(float s, float2 a, float2 c) {
    u64 rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(pk(s, s)), "l"(pk(a.x, a.y)), "l"(pk(c.x, c.y)));
    return up(rd);
}
// a * t = t.x (a.x, a.y) + t.y (-a.y, a.x)
__device__ __forceinline__ float2 cmul(float2 a, float2 t) {
    u64 m, rd;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(m) : "l"(pk(t.y, t.y)), "l"(pk(a.y, a.x)));
    const float2 mm = up(m);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(pk(t.x, t.x)), "l"(pk(a.x, a.y)), "l"(pk(-mm.x, mm.y)));
    return up(rd);
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// 8-point forward DFT, natural order in and out (decimation in frequency, 3 radix-2 levels).
// PRUNE: only outputs 0, 1, 6, 7 are produced (the last stage feeds bins 4..127 and 385..508 only).
template <bool PRUNE>
__device__ __forceinline__ void fft8(float2 (&x)[8]) {
    const float h = 0.70710678118654752440f;
    float2 t0 = cadd(x[0], x[4]), t4 = csub(x[0], x[4]);
    float2 t1 = cadd(x[1], x[5]), t5 = csub(x[1], x[5]);
    float2 t2 = cadd(x[2], x[6]), t6 = csub(x[2], x[6]);
    float2 t3 = cadd(x[3], x[7]), t7 = csub(x[3], x[7]);
    t5 = scale2(h, cadd(t5, mul_mi(t5)));                          // * W8^1 = h (1 - i)
    t6 = mul_mi(t6);                                               // * W8^2
    t7 = scale2(h, csub(mul_mi(t7), t7));                          // * W8^3 = -h (1 + i)
    float2 u0 = cadd(t0, t2), u2 = csub(t0, t2);
    float2 u1 = cadd(t1, t3), u3 = mul_mi(csub(t1, t3));
    float2 u4 = cadd(t4, t6), u6 = csub(t4, t6);
    float2 u5 = cadd(t5, t7), u7 = mul_mi(csub(t5, t7));
    x[0] = cadd(u0, u1);
    x[1] = cadd(u4, u5);
    x[6] = csub(u2, u3);
    x[7] = csub(u6, u7);
    if (!PRUNE) {
        x[4] = csub(u0, u1);
        x[2] = cadd(u2, u3);
        x[5] = csub(u4, u5);
        x[3] = cadd(u6, u7);
    }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Block-wide reduction over kFeatWarps warps; every thread gets the result.  `red` needs >= 16
// floats and is safe to reuse after the call returns (two barriers inside).
template <bool IS_MAX>
__device__ __forceinline__ float block_reduce(float v, float* red) {
    v = IS_MAX ? warp_max(v) : warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[lane & (kFeatWarps - 1)];
#pragma unroll
    for (int o = kFeatWarps / 2; o; o >>= 1) {
        float other = __shfl_xor_sync(0xffffffffu, r, o);
        r = IS_MAX ? fmaxf(r, other) : r + other;
    }
    __syncthreads();
    return r;
}

// Sum of a float2 over the block; every thread gets the result.  `red` needs >= 16 floats.
__device__ __forceinline__ float2 block_reduce_sum2(float2 v, float* red) {
    v.x = warp_sum(v.x);
    v.y = warp_sum(v.y);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[warp] = v.x;
        red[kFeatWarps + warp] = v.y;
    }
    __syncthreads();
    float a = red[lane & (kFeatWarps - 1)], b = red[kFeatWarps + (lane & (kFeatWarps - 1))];
#pragma unroll
    for (int o = kFeatWarps / 2; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    __syncthreads();
    return make_float2(a, b);
}

// AmplitudeToDB('power'): 10 log10(max(x, 1e-10)) (torchaudio functional.py:390) via MUFU.LG2
__device__ __forceinline__ float power_to_db(float x) {
    return 3.01029995663981195f * __log2f(fmaxf(x, 1e-10f));      // 10 * log10(2) * log2(x)
}

// za = Z[k], zb = Z[512 - k] of z = x1 + i x2: {|Z + conj Z'|^2, |Z - conj Z'|^2} = the two frames' power at bin k
// (the window carries the factor 1/2).  S = za + zb = (sr, di), D = za - zb = (dr, si).
__device__ __forceinline__ float2 pair_power(float2 za, float2 zb) {
    const float2 S = cadd(za, zb), D = csub(za, zb);
    u64 t, rd;
    asm("mul.f32x2 %0, %1, %1;" : "=l"(t) : "l"(pk(D.y, D.x)));                         // (si^2, dr^2)
    asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(rd) : "l"(pk(S.x, S.y)), "l"(t));          // (sr^2 + si^2, di^2 + dr^2)
    return up(rd);
}

// One windowed sample pair for FFT point n of the frame pair starting at sample `base0`.
// T = float (the reference's waveform type) or int16_t (PCM as stored in wav files: x / 32768 is exact in fp32 and the
// peak normalisation cancels the power-of-two scale, so both types give bit-identical features)
// PRE: the sample is the pre-emphasised one, y[i] = x[i] - coef * x[i-1], y[0] = x[0] (apply_pre_emphasis,
// src/preprocessing.py:213-239; product and difference rounded separately like the reference's tensor ops).  Only the
// edge pairs can touch i = 0; the shared stream frames never do (they start >= 120 samples into every window).
template <bool EDGE, bool PRE, class T>
__device__ __forceinline__ float sample_at(const T* __restrict__ clip, int i, float coef) {
    const float x = (float)clip[i];
    if (!PRE) return x;
    if (EDGE && i == 0) return x;
    return __fsub_rn(x, __fmul_rn(coef, (float)clip[i - 1]));
}
template <bool EDGE, bool PRE, class T>
__device__ __forceinline__ float2 load_point(const T* __restrict__ clip, float w, int n, int base0, bool second, float coef) {
    int i0 = base0 + n;
    int i1 = i0 + kHop;
    if (EDGE) {   // reflect padding of torch.stft(center=True, pad_mode="reflect")
        i0 = i0 < 0 ? -i0 : i0;
        i0 = i0 >= kSeg ? 2 * (kSeg - 1) - i0 : i0;
        i1 = i1 < 0 ? -i1 : i1;
        i1 = i1 >= kSeg ? 2 * (kSeg - 1) - i1 : i1;
        return scale2(w, make_float2(sample_at<EDGE, PRE, T>(clip, i0, coef), second ? sample_at<EDGE, PRE, T>(clip, i1, coef) : 0.f));
    }
    return scale2(w, make_float2(sample_at<EDGE, PRE, T>(clip, i0, coef), sample_at<EDGE, PRE, T>(clip, i1, coef)));
}

// The per-warp work item: frames 2p and 2p+1 -> 2 x 64 dB values in the tile.  Returns the lane's
// running max of the dB values it produced.
// RAW = false: dB values go to dst[band * band_stride + {0, frame_stride}] (the smem tile).
// RAW = true : un-logged mel POWER goes there instead (shared STFT frames of the streaming path).
template <bool EDGE, bool RAW, bool PRE, class SM, class T>
__device__ __forceinline__ float frame_pair(SM& S, const T* __restrict__ clip, float2* __restrict__ xb,
                                            const float (&wA)[7], const float (&wB)[6],
                                            const float2 (&mw)[kMelSlots], int bin_lo, int bin_hi, int base0, bool second,
                                            int lane, float* __restrict__ dst, int band_stride, int frame_stride,
                                            float coef = 0.f) {
    const int c = lane & 7, j = lane >> 3;             // base0: sample index of window tap 0 of the first frame

    // ---- stage 1: radix-8 over a (n = 64a + m), m = lane and lane + 32 ------------------------
    float2 A[8], B[8];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        A[a] = load_point<EDGE, PRE, T>(clip, wA[a], 64 * a + lane, base0, second, coef);
        B[a] = load_point<EDGE, PRE, T>(clip, wB[a], 64 * a + lane + 32, base0, second, coef);
    }
    A[6] = lane < 16 ? load_point<EDGE, PRE, T>(clip, wA[6], 384 + lane, base0, second, coef) : make_float2(0.f, 0.f);
    A[7] = make_float2(0.f, 0.f);                   // n >= 400: zero padding of the 512-point frame
    B[6] = make_float2(0.f, 0.f);
    B[7] = make_float2(0.f, 0.f);
    fft8<false>(A);
    fft8<false>(B);
    // twiddles W512^(m r): m = lane from the table, m = lane + 32 as table value * W16^r (a compile-time
    // constant) -- 4 FP instructions instead of a second shared-memory load per r
#pragma unroll
    for (int r = 1; r < 8; ++r) {
        const float2 tw = S.tw1[r - 1][lane];
        A[r] = cmul(A[r], tw);
        constexpr float kC[8] = {1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                                 0.f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
        constexpr float kS[8] = {0.f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
                                 -1.f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f};
        B[r] = cmul(B[r], cmul(tw, make_float2(kC[r], kS[r])));      // W16^r = exp(-2 pi i r / 16)
    }

    // ---- exchange 1 (two halves of r) + stage 2: radix-8 over b (m = 8b + c) -------------------
    float2 C[8], D[8];
#if CD_EXP == 2       // experiment: no exchanges at all
#pragma unroll
    for (int q = 0; q < 8; ++q) { C[q] = A[q]; D[q] = B[q]; }
#else
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        xb[r * kXbRow + lane] = A[r];
        xb[r * kXbRow + lane + 32] = B[r];
    }
    __syncwarp();
#pragma unroll
    for (int b = 0; b < 8; ++b) C[b] = xb[j * kXbRow + 8 * b + c];      // r = j
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        xb[r * kXbRow + lane] = A[r + 4];
        xb[r * kXbRow + lane + 32] = B[r + 4];
    }
    __syncwarp();
#pragma unroll
    for (int b = 0; b < 8; ++b) D[b] = xb[j * kXbRow + 8 * b + c];      // r = j + 4
    __syncwarp();
#endif
    fft8<false>(C);
    fft8<false>(D);
#pragma unroll
    for (int u = 1; u < 8; ++u) {
        const float2 t = S.tw2[u - 1][c];
        C[u] = cmul(C[u], t);
        D[u] = cmul(D[u], t);
    }

    // ---- exchange 2 + stage 3: radix-8 over c', pruned to s in {0,1,6,7} -----------------------
#if CD_EXP == 1 || CD_EXP == 2        // experiment: no second exchange (wrong results; measures what the exchange costs)
#pragma unroll
    for (int q = 0; q < 8; ++q) { A[q] = C[q]; B[q] = D[q]; }
#else
#pragma unroll
    for (int u = 0; u < 8; ++u) xb[j * kXbRow + 9 * u + c] = C[u];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) A[q] = xb[j * kXbRow + 9 * c + q];
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 8; ++u) xb[j * kXbRow + 9 * u + c] = D[u];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) B[q] = xb[j * kXbRow + 9 * c + q];
    __syncwarp();
#endif
    fft8<true>(A);      // A[s] = Z[64 s + 8 c + j]
    fft8<true>(B);      // B[s] = Z[64 s + 8 c + j + 4]

    // ---- unpack the two real spectra and take |.|^2 for bins < 128 ------------------------------
    // z = x1 + i x2  =>  Z[k] + conj Z[512-k] = 2 X1[k],  Z[k] - conj Z[512-k] = 2i X2[k].
    // Z[512-k] of my A[s] (k = 64s + 8c + j) sits in lane (7-c, 4-j) as B[7-s] (j != 0) or in lane
    // ((8-c)&7, 0) as A[7-s] (j == 0); of my B[s] (k + 4) in lane (7-c, 4-j) as A[7-s] (j != 0) or in
    // lane (7-c, 0) as B[7-s] (j == 0).  Each lane serves exactly one requester, so the sender selects.
    {
        const int srcA = j ? ((7 - c) | ((4 - j) << 3)) : ((8 - c) & 7);
        const int srcB = j ? ((7 - c) | ((4 - j) << 3)) : (7 - c);
        float2 pA[2], pB[2];
#pragma unroll
        for (int sidx = 0; sidx < 2; ++sidx) {
            const float2 sendA = j ? B[7 - sidx] : A[7 - sidx];
            const float2 sendB = j ? A[7 - sidx] : B[7 - sidx];
            pA[sidx].x = __shfl_sync(0xffffffffu, sendA.x, srcA);
            pA[sidx].y = __shfl_sync(0xffffffffu, sendA.y, srcA);
            pB[sidx].x = __shfl_sync(0xffffffffu, sendB.x, srcB);
            pB[sidx].y = __shfl_sync(0xffffffffu, sendB.y, srcB);
        }
        if (lane == 0) pA[1] = A[7];                 // k = 64: 512 - 64 = 448 = Z of this very lane (carry case)
        // float2 {P_frame0, P_frame1} at padded bin index pidx(k) = k + 2 (k / 16): the padding found (offline
        // search over the real band tables) to cost the fewest wavefronts for these stores + the mel reads
        const int o = 8 * c + 2 * (c >> 1) + j;
#pragma unroll
        for (int sidx = 0; sidx < 2; ++sidx) {
            xb[72 * sidx + o] = pair_power(A[sidx], pA[sidx]);
            xb[72 * sidx + o + 4] = pair_power(B[sidx], pB[sidx]);
        }
    }
    __syncwarp();

    // ---- banded mel contraction (cd_common.cuh): the lane reads the bins of its two segments once, for both
    //      frames, and hands the rising-side sums to the lane above ----------------------------------------
    float2 lo = make_float2(0.f, 0.f), hi = make_float2(0.f, 0.f), ulo = lo, uhi = lo;
    {
        // unused slots read the shared zero instead of being predicated off: one broadcast address in a bank pair
        // (8) the live ones rarely touch -- 19 wavefronts for the six loads, no predicates, no conditional moves
        const uint32_t xb_s = smem_u32(xb), zero_s = smem_u32(&S.zero2);
        const int b0 = bin_lo & 0xff, n0 = bin_lo >> 8;
#pragma unroll
        for (int s = 0; s < (CD_EXP == 5 ? 1 : kMelSlotsLo); ++s) {
            const int bin = b0 + s;
            const float2 pw = lds64(s < n0 ? xb_s + 8u * (uint32_t)(bin + 2 * (bin >> 4)) : zero_s);
            lo = fma2s(mw[s].x, pw, lo);
            ulo = fma2s(mw[s].y, pw, ulo);
        }
        const int b1 = bin_hi & 0xff, n1 = bin_hi >> 8;
#pragma unroll
        for (int s = 0; s < (CD_EXP == 5 ? 1 : kMelSlotsHi); ++s) {
            const int bin = b1 + s;
            const float2 pw = lds64(s < n1 ? xb_s + 8u * (uint32_t)(bin + 2 * (bin >> 4)) : zero_s);
            hi = fma2s(mw[kMelSlotsLo + s].x, pw, hi);
            uhi = fma2s(mw[kMelSlotsLo + s].y, pw, uhi);
        }
        // band m += up(m - 1): lanes 1..31 take both sums from lane - 1; lane 0's band 32 takes up(31) from lane 31's
        // lo segment (its band 0 has nothing below it, and up(63) is identically zero)
        const int below = (lane + 31) & 31;
        const float ax = __shfl_sync(0xffffffffu, ulo.x, below), ay = __shfl_sync(0xffffffffu, ulo.y, below);
        const float bx = __shfl_sync(0xffffffffu, uhi.x, below), by = __shfl_sync(0xffffffffu, uhi.y, below);
        if (lane) {
            lo.x += ax; lo.y += ay;
            hi.x += bx; hi.y += by;
        } else {
            hi.x += ax; hi.y += ay;
        }
    }
    __syncwarp();                                    // xb is rewritten by the next pair

    // ---- AmplitudeToDB: 10 log10(max(x, 1e-10)) (torchaudio functional.py:390) ------------------
    float* row_lo = dst + lane * band_stride;
    float* row_hi = dst + (lane + 32) * band_stride;
    if (RAW) {
        row_lo[0] = lo.x;
        row_hi[0] = hi.x;
        float m = fmaxf(lo.x, hi.x);
        if (second) {
            row_lo[frame_stride] = lo.y;
            row_hi[frame_stride] = hi.y;
            m = fmaxf(m, fmaxf(lo.y, hi.y));
        }
        return m;                                    // running max of the mel POWER (the PCEN variant's top_db needs it)
    }
    const float d0 = power_to_db(lo.x), d2 = power_to_db(hi.x);
    row_lo[0] = d0;
    row_hi[0] = d2;
    float m = fmaxf(d0, d2);
    if (second) {
        const float d1 = power_to_db(lo.y), d3 = power_to_db(hi.y);
        row_lo[frame_stride] = d1;
        row_hi[frame_stride] = d3;
        m = fmaxf(m, fmaxf(d1, d3));
    }
    return m;
}

// DCT-II of the floored dB tile for frame t, coefficients of one parity (PAR = 0: c = 0,2,..,12;
// PAR = 1: c = 1,3,..,11).  cos(pi (m + 1/2) c / 64) is even (c even) or odd (c odd) about m = 31.5, so
// each coefficient is a 32-term sum over d[m] +- d[63 - m], shared by all coefficients of that parity.
// POWER: the tile holds mel power (PCEN variant); dB and its top_db floor are applied on the way in.
// Returns {sum, sum of squares} of the coefficients it wrote (the block's mean / std come from one reduction).
template <int PAR, bool POWER>
__device__ __forceinline__ float2 mfcc_par(const float* __restrict__ db, float* __restrict__ mf, int t, float floor_db) {
    constexpr int NC = PAR ? 6 : 7;
    float acc[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[i] = 0.f;
#pragma unroll
    for (int m = 0; m < kMels / 2; ++m) {
        float lo = db[m * kFrames + t], hi = db[(kMels - 1 - m) * kFrames + t];
        if (POWER) {
            lo = fmaxf(power_to_db(lo), floor_db);
            hi = fmaxf(power_to_db(hi), floor_db);
        }
        const float d = PAR ? lo - hi : lo + hi;
#pragma unroll
        for (int i = 0; i < NC; ++i) acc[i] = fmaf(c_dct[(PAR + 2 * i) * kMels + m], d, acc[i]);
    }
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        mf[(PAR + 2 * i) * kFrames + t] = acc[i];
        s += acc[i];
        s2 = fmaf(acc[i], acc[i], s2);
    }
    return make_float2(s, s2);
}

// Everything after the tile is complete: top_db floor, mel rows, DCT-13, mean / unbiased std
// normalisation, deltas; fp32 features to `out` and/or the stem's bf16 map to `mo` (either may be null).
// PCEN = false: the tile holds dB, `top` is its block max; rows 0..63 are clamp((dB + 80) / 80).
// PCEN = true : the tile holds mel POWER, `top` is its block max; rows 0..63 are apply_pcen (src/preprocessing.py:
//               305-340) scaled to [0, 1] by the clip's min / max (:400-404); the MFCC rows still come from the dB.
// Ends with the block in sync except for the final stores.
// dd: also emit the delta-delta rows 90..102 = compute_deltas(compute_deltas(mfcc)) (:463-466).
template <bool PCEN, class SM>
__device__ __forceinline__ void feat_tail(SM& S, float top, float* __restrict__ out, __nv_bfloat16* __restrict__ mo, int tid,
                                          bool dd = false) {
    const float floor_db = (PCEN ? power_to_db(top) : top) - 80.0f;    // top_db = 80

    if (!PCEN) {
        // ---- mel block rows 0..63: max(dB, floor) -> (dB + 80) / 80 -> clamp ------------------------
        // The CNN stem's operand rides along: the same values rounded to bf16, scattered straight into the clip's
        // blocked map (cd_common.cuh: bmap_offset).  Its padding elements are never written by anyone: the map
        // lives in the tensor-core workspace, which is zeroed once at allocation.
        for (int i = tid; i < kMels * kFrames; i += kFeatThreads) {
            const float v = fmaxf(S.db[i], floor_db);
            S.db[i] = v;
            const float x01 = fminf(fmaxf((v + 80.0f) * 0.0125f, 0.f), 1.f);
            if (out) out[i] = x01;
            if (mo) {
                const int r = i / kFrames;
                mo[bmap_offset(r + 3, i - r * kFrames + 3)] = __float2bfloat16_rn(x01);
            }
        }
        __syncthreads();
    }

    // ---- MFCC = DCT-II(13 x 64) of the floored dB, then (x - mean) / (std + 1e-8) ---------------
    float* mf = reinterpret_cast<float*>(&S.xb[0][0]);             // [13][101], FFT scratch is idle now
    float2 part = make_float2(0.f, 0.f);
    {
        static_assert(kFeatThreads == 256, "two parity groups of 128 threads");
        const int t = tid & 127;
        if (t < kFrames) part = (tid >> 7) ? mfcc_par<1, PCEN>(S.db, mf, t, floor_db) : mfcc_par<0, PCEN>(S.db, mf, t, floor_db);
    }
    // mean and unbiased variance from one reduction of {sum, sum of squares}: |mean| is a fraction of the spread
    // here (c0 dominates both), so the one-pass form loses nothing in fp32 (checked against float64 on the goldens)
    constexpr int kN = kMfcc * kFrames;                            // 1313
    const float2 tot = block_reduce_sum2(part, S.red);
    const float mean = tot.x * (1.0f / kN);
    const float var = fmaxf(tot.y - tot.x * mean, 0.f) * (1.0f / (kN - 1));
    const float inv_std = 1.0f / (sqrtf(var) + 1e-8f);
    {
        for (int i = tid; i < kN; i += kFeatThreads) {
            const int cidx = i / kFrames, t = i - cidx * kFrames;
            const float* row = mf + cidx * kFrames;
            // __fmul_rn: never contracted into an FMA, so the bf16 map below sees identical values
            const float x0 = __fmul_rn(row[t] - mean, inv_std);
            const float xp = __fmul_rn(row[min(t + 1, kFrames - 1)] - mean, inv_std);
            const float xm = __fmul_rn(row[max(t - 1, 0)] - mean, inv_std);
            const float dl = (xp - xm) * 0.5f;                             // compute_deltas (:353-355)
            float d2 = 0.f;
            if (dd) {
                auto xn = [&](int u) { return __fmul_rn(row[min(max(u, 0), kFrames - 1)] - mean, inv_std); };
                auto d1 = [&](int u) { u = min(max(u, 0), kFrames - 1); return (xn(u + 1) - xn(u - 1)) * 0.5f; };
                d2 = (d1(t + 1) - d1(t - 1)) * 0.5f;
            }
            if (out) {
                out[(kMels + cidx) * kFrames + t] = x0;
                out[(kMels + kMfcc + cidx) * kFrames + t] = dl;
                if (dd) out[(kMels + 2 * kMfcc + cidx) * kFrames + t] = d2;
            }
            if (mo) {
                mo[bmap_offset(kMels + cidx + 3, t + 3)] = __float2bfloat16_rn(x0);
                mo[bmap_offset(kMels + kMfcc + cidx + 3, t + 3)] = __float2bfloat16_rn(dl);
                if (dd) mo[bmap_offset(kMels + 2 * kMfcc + cidx + 3, t + 3)] = __float2bfloat16_rn(d2);
            }
        }
    }

    if (PCEN) {
        // ---- PCEN rows 0..63 from the power tile: smoother = zero-padded mean of frames t-5 .. t+4 (avg_pool2d
        //      kernel (1, 10), padding (0, 5), cut to 101 columns); (P / (eps + M)^0.98 + 2)^0.5 - 2^0.5 ------------
        constexpr int kPer = (kMels * kFrames + kFeatThreads - 1) / kFeatThreads;      // 26
        float val[kPer];
        float vmin = INFINITY, vmax = -INFINITY;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int i = tid + k * kFeatThreads;
            float v = 0.f;
            if (i < kMels * kFrames) {
                const int r = i / kFrames, t = i - r * kFrames;
                const float* row = S.db + r * kFrames;
                float sm = 0.f;
#pragma unroll
                for (int j = -5; j <= 4; ++j) sm += (t + j >= 0 && t + j < kFrames) ? row[t + j] : 0.f;
                sm *= 0.1f;
                v = sqrtf(__fdiv_rn(row[t], powf(1e-6f + sm, 0.98f)) + 2.0f) - 1.41421354f;
                vmin = fminf(vmin, v);
                vmax = fmaxf(vmax, v);
            }
            val[k] = v;
        }
        vmax = block_reduce<true>(vmax, S.red);
        vmin = -block_reduce<true>(-vmin, S.red);
        const float range = (vmax - vmin) + 1e-8f;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int i = tid + k * kFeatThreads;
            if (i < kMels * kFrames) {
                const float x01 = __fdiv_rn(val[k] - vmin, range);
                if (out) out[i] = x01;
                if (mo) {
                    const int r = i / kFrames;
                    mo[bmap_offset(r + 3, i - r * kFrames + 3)] = __float2bfloat16_rn(x01);
                }
            }
        }
    }
}

template <class T>
struct ClipDesc {
    const T* src;         // first sample of the clip
    int len;              // samples available
};

template <class T>
__device__ __forceinline__ ClipDesc<T> clip_desc(const T* audio, const int64_t* offs, const int32_t* lens,
                                              int64_t stride_or_hop, int mode, int64_t b) {
    ClipDesc<T> d;
    if ((mode & kModeMask) == kModeVar) {
        d.src = audio + offs[b];
        d.len = lens[b];
    } else {
        d.src = audio + b * stride_or_hop;
        d.len = kSeg;
    }
    return d;
}
template <class T>
__device__ __forceinline__ bool bulk_ok(const ClipDesc<T>& d) {
    return d.len == kSeg && ((reinterpret_cast<uintptr_t>(d.src) & 15) == 0);
}

// PRE / PCEN: the use_pre_emphasis / use_pcen variants of the 90-feature path (the part that keeps
// the CNN input at 90 rows); <false, false> is the training configuration (src/train.py:264-287) and the benchmark.
template <class T, bool PRE, bool PCEN>
__global__ void __launch_bounds__(kFeatThreads, 2)
feat_kernel(const T* __restrict__ audio, const int64_t* __restrict__ offs, const int32_t* __restrict__ lens,
            int64_t stride_or_hop, int mode, int64_t B, float* __restrict__ feat,
            __nv_bfloat16* __restrict__ map_out, const FeatTables* __restrict__ tables, float coef) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FeatSmem& S = *reinterpret_cast<FeatSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- one-time per CTA: twiddles -> smem, per-lane mel taps -> registers, barrier ----------------
    {
        const float* src = reinterpret_cast<const float*>(tables);
        float* dst = reinterpret_cast<float*>(&S.tw1[0][0]);
        constexpr int n = (sizeof(S.tw1) + sizeof(S.tw2)) / 4;
        static_assert(offsetof(FeatTables, tw2) == sizeof(S.tw1) && offsetof(FeatSmem, tw2) - offsetof(FeatSmem, tw1) == sizeof(S.tw1), "layout");
        for (int i = tid; i < n; i += kFeatThreads) dst[i] = src[i];
        if (tid == 0) S.zero2 = make_float2(0.f, 0.f);
    }
    float2 mw[kMelSlots];
#pragma unroll
    for (int s = 0; s < kMelSlots; ++s) mw[s] = tables->melt[s][lane];
    const int bin_lo = tables->melb[0][lane], bin_hi = tables->melb[1][lane];
    if (tid == 0) {
        mbar_init(&S.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t phase = 0u;
    T* clip = reinterpret_cast<T*>(S.clip);
    constexpr uint32_t kClipBytes = kSeg * sizeof(T);
    // One bulk copy (TMA engine) brings a clip in.  It is issued as soon as the previous clip's samples are dead
    // (after its last FFT), so it lands under that clip's tail; the clip after that is pulled into L2 meanwhile.
    auto fetch = [&](int64_t bf) {                                  // thread 0 only
        if (bf < B) {
            const ClipDesc<T> df = clip_desc(audio, offs, lens, stride_or_hop, mode, bf);
            if (bulk_ok(df)) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses -> async proxy
                mbar_expect_tx(&S.bar, kClipBytes);
                bulk_g2s(clip, df.src, kClipBytes, &S.bar);
            }
        }
        const int64_t bn = bf + gridDim.x;
        if (bn < B) {
            const ClipDesc<T> dn = clip_desc(audio, offs, lens, stride_or_hop, mode, bn);
            if (bulk_ok(dn))
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(dn.src), "r"(kClipBytes) : "memory");
        }
    };
    if (tid == 0) fetch(blockIdx.x);
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        const ClipDesc<T> d = clip_desc(audio, offs, lens, stride_or_hop, mode, b);

        // ---- stage the clip + peak ------------------------------------------------------------------
        float peak = 0.f;
        if (bulk_ok(d)) {
            mbar_wait(&S.bar, phase);
            phase ^= 1u;
            if (sizeof(T) == 4) {
                const float4* c4 = reinterpret_cast<const float4*>(clip);
                for (int i = tid; i < kSeg / 4; i += kFeatThreads) {
                    const float4 v = c4[i];
                    peak = fmaxf(peak, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
                }
            } else {
                const short4* c4 = reinterpret_cast<const short4*>(clip);
                for (int i = tid; i < kSeg / 4; i += kFeatThreads) {
                    const short4 v = c4[i];
                    peak = fmaxf(peak, fmaxf(fmaxf(fabsf((float)v.x), fabsf((float)v.y)), fmaxf(fabsf((float)v.z), fabsf((float)v.w))));
                }
            }
        } else {
            // AudioPreprocessor.process: normalise over the WHOLE clip, then centre-trim / zero-pad
            const int n = d.len;
            for (int i = tid; i < n; i += kFeatThreads) peak = fmaxf(peak, fabsf((float)d.src[i]));
            const int start = n > kSeg ? (n - kSeg) / 2 : 0;          // trim offset (src/preprocessing.py:378)
            const int padl = n < kSeg ? (kSeg - n) / 2 : 0;           // left zero pad (:383)
            for (int i = tid; i < kSeg; i += kFeatThreads) {
                const int sidx = i - padl;
                clip[i] = (sidx >= 0 && sidx < n) ? d.src[start + sidx] : (T)0;
            }
        }
        peak = block_reduce<true>(peak, S.red);                        // its barriers also order the generic stores
        // normalize() (:199-212): x / max|x|, skipped for silence.  The scale is folded into this lane's
        // 13 window weights (held in registers for the whole clip) instead of rewriting the clip.
        const float inv = (peak > 0.f && !(mode & kModeNoNormalize)) ? 1.0f / peak : 1.0f;
        float wA[7], wB[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            wA[a] = __ldg(&tables->win[64 * a + lane]) * inv;
            wB[a] = __ldg(&tables->win[64 * a + lane + 32]) * inv;
        }
        wA[6] = lane < 16 ? __ldg(&tables->win[384 + lane]) * inv : 0.f;

        // ---- STFT -> mel -> dB, one frame pair per warp per round -----------------------------------
        float wmax = -INFINITY;
        for (int p = warp; p < (CD_EXP == 3 ? 0 : kPairs); p += kFeatWarps) {
            const bool edge = (p == 0) || (p >= 49);
            float* dst = S.db + 2 * p;
            const int base0 = 2 * kHop * p - kWin / 2;
            const bool second = (2 * p + 1) < kFrames;
            wmax = fmaxf(wmax, edge ? frame_pair<true, PCEN, PRE>(S, clip, S.xb[warp], wA, wB, mw, bin_lo, bin_hi, base0, second, lane, dst, kFrames, 1, coef)
                                    : frame_pair<false, PCEN, PRE>(S, clip, S.xb[warp], wA, wB, mw, bin_lo, bin_hi, base0, second, lane, dst, kFrames, 1, coef));
        }
        const float top = block_reduce<true>(wmax, S.red);             // barrier: dB tile complete, clip buffer dead
        if (tid == 0) fetch(b + gridDim.x);
#if CD_EXP == 4       // experiment: no tail (keeps one store per thread so nothing upstream is dead code)
        if (feat) feat[b * (int64_t)kFeatElem + tid] = top + S.db[tid * 25];
        if (map_out) map_out[b * (int64_t)kMapElems + tid] = __float2bfloat16_rn(top + S.db[tid * 25]);
#else
        const bool dd = (mode & kModeDeltaDelta) != 0;
        feat_tail<PCEN>(S, top, feat ? feat + b * (int64_t)(dd ? kFeatElemDD : kFeatElem) : nullptr,
                        map_out ? map_out + b * (int64_t)(dd ? NetShape<kFeatDD>::MAP_ELEMS : kMapElems) : nullptr, tid, dd);
#endif
        __syncthreads();                                               // xb / db free again
    }
}

// ----------------------------------------------------------------------------------------------
// streaming with shared STFT frames (RealtimePreprocessor.add_audio, src/preprocessing.py:582-612)
// ----------------------------------------------------------------------------------------------
// With hop = 160 h, window j's frame t is the stream's global frame j h + t (centred at sample
// 160 (j h + t)); only frames 0, 1, 99, 100 of a window see its reflect padding.  So the mel POWER of
// every global frame is computed once (stream_mel_kernel), each window takes frames 2..97 from there,
// scaled by 1/peak^2 (peak normalisation is a scalar on the samples), and recomputes just the pairs
// (0,1), (98,99), (100,-) from its own samples.  The window peak is the max of 100 per-hop block maxima.
struct MelSmem {
    float2 xb[kFeatWarps][kXbElems];
    float2 tw1[7][64];
    float2 tw2[7][8];
    float2 zero2;
};
static_assert(offsetof(MelSmem, xb) % 128 == 0 && offsetof(MelSmem, zero2) % 128 == 64, "zero2 sits in bank pair 8 of xb's numbering");
struct StreamSmem {
    float2 xb[kFeatWarps][kXbElems];
    float  db[kMels * kFrames];
    float2 tw1[7][64];
    float2 tw2[7][8];
    float2 zero2;
    float  red[64];
};
static_assert(offsetof(StreamSmem, xb) % 128 == 0 && offsetof(StreamSmem, zero2) % 128 == 64, "zero2 sits in bank pair 8 of xb's numbering");

template <class SM>
__device__ __forceinline__ void load_lane_tables(SM& S, const FeatTables* __restrict__ tables, float2 (&mw)[kMelSlots],
                                                 int& bin_lo, int& bin_hi, int tid, int lane) {
    const float* src = reinterpret_cast<const float*>(tables);
    float* dst = reinterpret_cast<float*>(&S.tw1[0][0]);
    constexpr int n = (sizeof(S.tw1) + sizeof(S.tw2)) / 4;
    for (int i = tid; i < n; i += kFeatThreads) dst[i] = src[i];
    if (tid == 0) S.zero2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int s = 0; s < kMelSlots; ++s) mw[s] = tables->melt[s][lane];
    bin_lo = tables->melb[0][lane];
    bin_hi = tables->melb[1][lane];
}

__device__ __forceinline__ void lane_window(const FeatTables* __restrict__ tables, float inv, int lane, float (&wA)[7], float (&wB)[6]) {
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        wA[a] = __ldg(&tables->win[64 * a + lane]) * inv;
        wB[a] = __ldg(&tables->win[64 * a + lane + 32]) * inv;
    }
    wA[6] = lane < 16 ? __ldg(&tables->win[384 + lane]) * inv : 0.f;
}

// bmax[i] = max |x[160 i .. 160 i + 159]|, one warp per block
__global__ void __launch_bounds__(256)
block_max_kernel(const float* __restrict__ audio, int64_t n_blocks, float* __restrict__ bmax) {
    const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n_blocks) return;
    const float* p = audio + i * kHop;
    float m = 0.f;
    for (int k = threadIdx.x & 31; k < kHop; k += 32) m = fmaxf(m, fabsf(p[k]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) bmax[i] = m;
}

// melp[f][64] = un-normalised mel power of global frame f, f in [2, F)
template <bool PRE>
__global__ void __launch_bounds__(kFeatThreads, 2)
stream_mel_kernel(const float* __restrict__ audio, int64_t F, float* __restrict__ melp, const FeatTables* __restrict__ tables,
                  float coef) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MelSmem& S = *reinterpret_cast<MelSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2 mw[kMelSlots];
    float wA[7], wB[6];
    int bin_lo, bin_hi;
    load_lane_tables(S, tables, mw, bin_lo, bin_hi, tid, lane);
    lane_window(tables, 1.0f, lane, wA, wB);
    __syncthreads();
    const int64_t n_pairs = (F - 2 + 1) / 2;
    for (int64_t q = (int64_t)blockIdx.x * kFeatWarps + warp; q < n_pairs; q += (int64_t)gridDim.x * kFeatWarps) {
        int64_t f0 = 2 + 2 * q;
        if (f0 + 1 >= F) f0 = F - 2;                  // odd tail: redo (F-2, F-1); both writers store identical values
        const float* base = audio + f0 * kHop - kWin / 2;
        frame_pair<false, true, PRE>(S, base, S.xb[warp], wA, wB, mw, bin_lo, bin_hi, 0, true, lane, melp + f0 * kMels, 1, kMels, coef);
    }
}

template <bool PRE, bool PCEN>
__global__ void __launch_bounds__(kFeatThreads, 2)
stream_window_kernel(const float* __restrict__ audio, int64_t hop, int64_t n_windows, const float* __restrict__ melp,
                     const float* __restrict__ bmax, float* __restrict__ feat, __nv_bfloat16* __restrict__ map_out,
                     const FeatTables* __restrict__ tables, float coef, int dd) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StreamSmem& S = *reinterpret_cast<StreamSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2 mw[kMelSlots];
    int bin_lo, bin_hi;
    load_lane_tables(S, tables, mw, bin_lo, bin_hi, tid, lane);
    __syncthreads();
    const int64_t h = hop / kHop;
    for (int64_t j = blockIdx.x; j < n_windows; j += gridDim.x) {
        float peak = tid < kSeg / kHop ? bmax[j * h + tid] : 0.f;
        peak = block_reduce<true>(peak, S.red);
        const float inv = peak > 0.f ? 1.0f / peak : 1.0f;           // normalize(): skipped for silence
        float wmax = -INFINITY;
        if (warp < 3) {
            // frames that touch this window's reflect padding: pairs (0,1), (98,99), (100,-)
            float wA[7], wB[6];
            lane_window(tables, inv, lane, wA, wB);
            const int p = warp == 0 ? 0 : 48 + warp;
            wmax = frame_pair<true, PCEN, PRE>(S, audio + j * hop, S.xb[warp], wA, wB, mw, bin_lo, bin_hi, 2 * kHop * p - kWin / 2,
                                               (2 * p + 1) < kFrames, lane, S.db + 2 * p, kFrames, 1, coef);
        } else {
            // frames 2..97 come from the shared global frames j h + t
            const float inv2 = inv * inv;
            const float* src = melp + (j * h + 2) * kMels;
            for (int i = tid - 96; i < 96 * kMels; i += kFeatThreads - 96) {
                const int t = i >> 6, m = i & 63;
                const float d = PCEN ? src[i] * inv2 : power_to_db(src[i] * inv2);     // PCEN keeps the tile in power
                S.db[m * kFrames + t + 2] = d;
                wmax = fmaxf(wmax, d);
            }
        }
        feat_tail<PCEN>(S, block_reduce<true>(wmax, S.red), feat ? feat + j * (int64_t)(dd ? kFeatElemDD : kFeatElem) : nullptr,
                        map_out ? map_out + j * (int64_t)(dd ? NetShape<kFeatDD>::MAP_ELEMS : kMapElems) : nullptr, tid, dd != 0);
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------------------------
// host side: tables + launcher
// ----------------------------------------------------------------------------------------------
static float bits2f(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}

int build_feat_tables(cd_handle* h) {
    FeatTables* t = new FeatTables();
    const double two_pi = 6.283185307179586476925286766559;
    for (int r = 1; r < 8; ++r)
        for (int m = 0; m < 64; ++m) {
            const double a = -two_pi * (double)((m * r) % 512) / 512.0;
            t->tw1[r - 1][m] = make_float2((float)cos(a), (float)sin(a));
        }
    for (int u = 1; u < 8; ++u)
        for (int c = 0; c < 8; ++c) {
            const double a = -two_pi * (double)((c * u) % 64) / 64.0;
            t->tw2[u - 1][c] = make_float2((float)cos(a), (float)sin(a));
        }
    for (int n = 0; n < kWin; ++n) t->win[n] = 0.5f * bits2f(CD_HANN_BITS[n]);   // exact (power of two)
    {
        // bin -> the (at most two, adjacent) bands it feeds; segment m = bins whose lower band is m
        float w_of[64][129];
        int lower[129];
        memset(w_of, 0, sizeof(w_of));
        for (int k = 0; k < 129; ++k) lower[k] = -1;
        for (int m = 0; m < 64; ++m)
            for (int tp = 0; tp < CD_MEL_COUNT[m]; ++tp) {
                const int k = CD_MEL_START[m] + tp;
                w_of[m][k] = bits2f(CD_MEL_W_BITS[m * CD_MEL_MAX_TAPS + tp]);
                if (lower[k] < 0) lower[k] = m;
                else if (m != lower[k] + 1) { delete t; return set_err(h, CD_E_STATE, "mel bands are not adjacent triangles"); }
            }
        for (int lane = 0; lane < 32; ++lane)
            for (int half = 0; half < 2; ++half) {
                const int m = lane + 32 * half, slots = half ? kMelSlotsHi : kMelSlotsLo, s0 = half ? kMelSlotsLo : 0;
                int first = -1, n = 0;
                for (int k = 0; k < 129; ++k)
                    if (lower[k] == m) {
                        if (first < 0) first = k;
                        if (k != first + n || n >= slots || k > 127) { delete t; return set_err(h, CD_E_STATE, "mel table does not fit the slot layout"); }
                        t->melt[s0 + n][lane] = make_float2(w_of[m][k], m < 63 ? w_of[m + 1][k] : 0.f);
                        ++n;
                    }
                for (int s = n; s < slots; ++s) t->melt[s0 + s][lane] = make_float2(0.f, 0.f);
                t->melb[half][lane] = (first < 0 ? 4 : first) | (n << 8);
            }
    }
    cudaError_t e = cudaMalloc(&h->d_tables, sizeof(FeatTables));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_tables, t, sizeof(FeatTables), cudaMemcpyHostToDevice);
    float dct[kMfcc * kMels];
    for (int i = 0; i < kMfcc * kMels; ++i) dct[i] = bits2f(CD_DCT_BITS[i]);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_dct, dct, sizeof(dct));
    auto big = [&](auto kernel, size_t bytes) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    };
    big(feat_kernel<float, false, false>, sizeof(FeatSmem));
    big(feat_kernel<float, true, false>, sizeof(FeatSmem));
    big(feat_kernel<float, false, true>, sizeof(FeatSmem));
    big(feat_kernel<float, true, true>, sizeof(FeatSmem));
    big(feat_kernel<int16_t, false, false>, sizeof(FeatSmem));
    big(feat_kernel<int16_t, true, false>, sizeof(FeatSmem));
    big(feat_kernel<int16_t, false, true>, sizeof(FeatSmem));
    big(feat_kernel<int16_t, true, true>, sizeof(FeatSmem));
    big(stream_window_kernel<false, false>, sizeof(StreamSmem));
    big(stream_window_kernel<true, false>, sizeof(StreamSmem));
    big(stream_window_kernel<false, true>, sizeof(StreamSmem));
    big(stream_window_kernel<true, true>, sizeof(StreamSmem));
    delete t;
    if (e != cudaSuccess) return set_err(h, CD_E_CUDA, "feature tables upload", e);
    return CD_OK;
}

// the four (pre-emphasis, PCEN) variants of a kernel template, picked by the handle's options
#define CD_FEAT_VARIANT(pre, pcen, CALL)           \
    do {                                           \
        if (pre) {                                 \
            if (pcen) { CALL(true, true); }        \
            else      { CALL(true, false); }       \
        } else {                                   \
            if (pcen) { CALL(false, true); }       \
            else      { CALL(false, false); }      \
        }                                          \
    } while (0)

int launch_features(cd_handle* h, const float* d_audio, const int64_t* d_offsets, const int32_t* d_lengths,
                    int64_t stride_or_hop, int mode, float coef, int64_t B, float* d_feat, __nv_bfloat16* d_map,
                    cudaStream_t s) {
    if (B == 0) return CD_OK;
    const bool pre = (mode & kModePreEmph) != 0, pcen = (mode & kModePcen) != 0;
    const int64_t slots = 2 * (int64_t)h->sm_count;     // two resident CTAs per SM
    if ((mode & kModeMask) == kModeStream && stride_or_hop % kHop == 0 && stride_or_hop <= kSeg / 2 && B >= 2) {
        // overlapping windows on the frame grid: share the STFT frames
        const int64_t hop = stride_or_hop, hh = hop / kHop;
        const int64_t F = (B - 1) * hh + 99;                         // global frames 2 .. F-1 are needed
        const int64_t n_blocks = (B - 1) * hh + kSeg / kHop;
        const size_t need = ((size_t)F * kMels + (size_t)n_blocks) * sizeof(float);
        if (need > h->ws_stream_bytes) {
            CD_CUDA(h, cudaDeviceSynchronize());
            cudaFree(h->d_ws_stream);
            h->d_ws_stream = nullptr;
            h->ws_stream_bytes = 0;
            cudaError_t e = cudaMalloc(&h->d_ws_stream, need);
            if (e != cudaSuccess) return set_err(h, CD_E_NOMEM, "stream workspace allocation", e);
            h->ws_stream_bytes = need;
        }
        float* melp = h->d_ws_stream;
        float* bmax = melp + (size_t)F * kMels;
        block_max_kernel<<<(unsigned)((n_blocks + 7) / 8), 256, 0, s>>>(d_audio, n_blocks, bmax);
        const int64_t pair_ctas = ((F - 1) / 2 + kFeatWarps - 1) / kFeatWarps;
        const int mel_grid = (int)(pair_ctas < 4 * slots ? pair_ctas : 4 * slots);
        if (pre)
            stream_mel_kernel<true><<<mel_grid, kFeatThreads, sizeof(MelSmem), s>>>(d_audio, F, melp, h->d_tables, coef);
        else
            stream_mel_kernel<false><<<mel_grid, kFeatThreads, sizeof(MelSmem), s>>>(d_audio, F, melp, h->d_tables, coef);
#define CD_CALL(P, Q)                                                                                   \
    stream_window_kernel<P, Q><<<(int)(B < slots ? B : slots), kFeatThreads, sizeof(StreamSmem), s>>>(   \
        d_audio, hop, B, melp, bmax, d_feat, d_map, h->d_tables, coef, (mode & kModeDeltaDelta) ? 1 : 0)
        CD_FEAT_VARIANT(pre, pcen, CD_CALL);
#undef CD_CALL
        h->launches += 3;
        CD_CUDA(h, cudaGetLastError());
        return CD_OK;
    }
    const int grid = (int)(B < slots ? B : slots);
#define CD_CALL(P, Q)                                                                                          \
    feat_kernel<float, P, Q><<<grid, kFeatThreads, sizeof(FeatSmem), s>>>(d_audio, d_offsets, d_lengths, stride_or_hop, \
                                                                          mode, B, d_feat, d_map, h->d_tables, coef)
    CD_FEAT_VARIANT(pre, pcen, CD_CALL);
#undef CD_CALL
    h->launches += 1;
    CD_CUDA(h, cudaGetLastError());
    return CD_OK;
}

// int16 PCM clips (fixed length): same kernel, 32 000-byte bulk copies
int launch_features_i16(cd_handle* h, const int16_t* d_audio, int64_t stride, int mode, float coef, int64_t B,
                        float* d_feat, __nv_bfloat16* d_map, cudaStream_t s) {
    if (B == 0) return CD_OK;
    const int64_t slots = 2 * (int64_t)h->sm_count;
    const bool pre = (mode & kModePreEmph) != 0, pcen = (mode & kModePcen) != 0;
#define CD_CALL(P, Q)                                                                                            \
    feat_kernel<int16_t, P, Q><<<(int)(B < slots ? B : slots), kFeatThreads, sizeof(FeatSmem), s>>>(              \
        d_audio, nullptr, nullptr, stride, mode, B, d_feat, d_map, h->d_tables, coef)
    CD_FEAT_VARIANT(pre, pcen, CD_CALL);
#undef CD_CALL
    h->launches += 1;
    CD_CUDA(h, cudaGetLastError());
    return CD_OK;
}

void get_host_tables(float* hann, float* fb, float* dct) {
    if (hann)
        for (int n = 0; n < kWin; ++n) hann[n] = bits2f(CD_HANN_BITS[n]);
    if (fb) {
        memset(fb, 0, sizeof(float) * 257 * 64);
        for (int m = 0; m < 64; ++m)
            for (int tp = 0; tp < CD_MEL_COUNT[m]; ++tp)
                fb[(CD_MEL_START[m] + tp) * 64 + m] = bits2f(CD_MEL_W_BITS[m * CD_MEL_MAX_TAPS + tp]);
    }
    if (dct)
        for (int m = 0; m < 64; ++m)
            for (int c = 0; c < 13; ++c) dct[m * 13 + c] = bits2f(CD_DCT_BITS[c * 64 + m]);
}

}  // namespace cd
