// Fused feature kernel: clip -> (90 | 103, 101) features, persistent, four 128-thread CTAs per SM.
//
// Replaces, per clip, AudioPreprocessor.normalize (src/preprocessing.py:199-212), pad_or_trim
// (:358-385) and extract_features (:432-489): T.MelSpectrogram (reflect pad, 400-sample periodic
// Hann centred in a 512-point frame, rFFT, power, 64-band HTK mel), AmplitudeToDB(top_db=80),
// (dB+80)/80 clamp, T.MFCC's DCT-II of the floored dB, global mean / unbiased-std normalisation,
// and the replicate-padded central-difference deltas.  The reference computes the mel
// spectrogram twice (extract_mel_spectrogram and again inside T.MFCC); here it is computed once.
//
// What bounds it (round-1 ncu + tools/ubench.cu): the shared-memory data pipe (128 B/clk/SM, shuffles included)
// and the FMA pipe (packed f32x2 instructions issue at half rate, so they save issue slots, not FMA cycles).
// The design therefore minimises shared-memory wavefronts per clip:
//   * the clip is NOT staged in shared memory: lane m of a warp reads samples s0 + 32 q + m straight from
//     global memory (one coalesced 128-byte line per load) and a warp walks CONSECUTIVE frame pairs, so 8 of the
//     18 sample registers of a pair are carried over from the previous pair (10 loads per pair instead of 26);
//   * two frames ride in one complex 512-point FFT, z = w (x_2p + i x_2p+1), decomposed 16 x 16 x 2:
//     a radix-16 DFT in registers (n = 32 a + m), twiddles, ONE exchange through shared memory (16 STS.64 +
//     8 LDS.128, conflict-free), a second radix-16 DFT in registers, and a final radix-2 across lane pairs
//     by shuffle, pruned to the outputs that feed bins 4..127 (tools/fft_lane_model.py is the numpy model of
//     this index arithmetic);
//   * peak normalisation never touches the samples: the peak is tracked while the samples are loaded and
//     1 / peak^2 is applied to the mel POWER in the tail (exact for power-of-two gains; clips whose peak is
//     outside 2^+-24 are redone with a power-of-two pre-scale so nothing under- or overflows);
//   * the 64 x 101 tile holds mel power; scaling, dB, top_db floor, DCT, mean / std, deltas happen in the tail.
// The power spectrum is invariant to a circular shift of the frame, so the 400 windowed samples
// sit at n = 0..399 of the 512-point frame instead of torch.stft's 56..455.
#include "cd_common.cuh"
#include "cd_tables.h"
#include <stdlib.h>

namespace cd {

// Occupancy configuration (make exp builds the alternatives: tools/feat_time.py compares them):
//   CD_FEAT_WARPS x CD_FEAT_CTAS = warps per CTA x CTAs per SM.  Registers per thread follow from the product
//   (16 warps: 128, 21: 96, 24: 80); the FFT alone holds 32, the lane's window 13.
#ifndef CD_FEAT_WARPS
#define CD_FEAT_WARPS 4
#endif
#ifndef CD_FEAT_CTAS
#define CD_FEAT_CTAS 4
#endif
// CD_TW: where the lane's 15 stage-1 twiddles W512^(m r) come from.
//   0: six of them (r = 1, 2, 3, 4, 8, 12) in registers, the other nine as products (9 more complex multiplies per pair,
//      and twice the twiddle error: a product of two rounded twiddles is off by ~1.2 ulp, measured 2x on the PCEN rows);
//   2: all fifteen EXACT from a per-CTA shared-memory table, 8 LDS.128 per pair (32 wavefronts; the shared-memory pipe
//      has the headroom since the clip is no longer staged there) -- no twiddle registers at all.
#ifndef CD_TW
#define CD_TW 2
#endif
// The next pair's 10 new samples are requested before the current pair's FFT (requesting them after the FFT or after the mel
// contraction was measured slower: the loads then arrive too late to hide under anything).
// CD_L2PF: when the clip after this one is pulled into L2.  0: never;  1: when this clip starts;  2: when this clip's tail
//   starts (a shorter stay in L2: with 4 x 148 clips in flight a whole-clip lead is ~38 MB of prefetched lines on top of
//   the ~59 MB that stream through meanwhile);  3: each warp pulls its quarter half-way through its run.
//   Measured over 65 536 clips (ncu dram__bytes_read, CUDA-event time): 1: 6.54 GB, 52.9 ns/clip -- a third of the prefetched
//   lines are evicted before they are used and fetched again;  0: 4.19 GB (exactly the audio), 52.5;  2: 4.25 GB, 52.2;
//   3: 5.21 GB, 55.5.
#ifndef CD_L2PF
#define CD_L2PF 2
#endif
// CD_STCS: output stores carry the streaming (evict-first) hint
#ifndef CD_STCS
#define CD_STCS 1
#endif
// CD_ABLATE: timing-only builds (`make ablations`; their RESULTS ARE WRONG by construction, only the time matters).
//   1: no tail;  2: no mel contraction;  3: no FFT and no mel (loads, window, peak and tail only);  4: FFT without its exchange.
#ifndef CD_ABLATE
#define CD_ABLATE 0
#endif
#ifndef CD_PAIR_UNROLL
#define CD_PAIR_UNROLL 1
#endif
constexpr int kPairUnroll = CD_PAIR_UNROLL;      // unroll of the frame-pair loop (2 was measured: see profiles/r2_feat_ablation.txt)
#ifndef CD_PIN_CTX
#define CD_PIN_CTX 1
#endif

constexpr int kFeatWarps   = CD_FEAT_WARPS;
constexpr int kFeatThreads = kFeatWarps * 32;
constexpr int kFeatCtasPerSm = CD_FEAT_CTAS;  // one CTA's reductions / tail overlap the others' FFTs
constexpr int kXchRow      = 34;              // float2 per row of the exchange tile (16 rows: r = k mod 16)
constexpr int kXchElems    = 16 * kXchRow;    // 544
constexpr int kPowElems    = 160;             // the pair's power spectra at padded bin index < 146, zero slot at 152
constexpr int kPowZero     = 152;
constexpr int kXbElems     = kXchElems + kPowElems;   // 704 float2 = 5632 B per warp
constexpr int kTwPitch     = 18;              // float2 per lane row of the twiddle table: 16 used, 144-byte pitch
constexpr int kTilePitch   = 102;             // floats per band row of the power tile: even, so a frame PAIR is one 8-byte access
static_assert(kFeatWarps >= 4 && kFeatWarps <= 8, "tail: two DCT groups of two warps; block reductions over <= 8 warps");

__constant__ float c_dct[kMfcc * kMels];      // [13][64], ortho DCT-II rows

struct FeatSmem {
    float2 xb[kFeatWarps][kXbElems];          //  5 632 B per warp: FFT exchange + power spectra; MFCC scratch in the tail
    float  db[kMels * kTilePitch];            //  26 112 B, mel POWER tile [band][frame] (dB after the tail's first pass)
    float2 twl[32][kTwPitch];                 //  4 608 B, W512^(m r) per lane m, rows 144 B apart (conflict-free LDS.128)
    float  red[64];
};
static_assert(kFeatCtasPerSm * (sizeof(FeatSmem) + 1024) <= 227 * 1024, "CTAs per SM");
static_assert(sizeof(float2) * kXbElems % 128 == 0 && kPowZero % 16 == 8, "zero slot sits in bank pair 8 of the power region");
static_assert(sizeof(float) * kMfcc * kFrames <= sizeof(float2) * kXbElems, "MFCC scratch fits one warp's region");

// ----------------------------------------------------------------------------------------------
// small PTX wrappers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}

// Packed fp32 pairs (add/sub/mul/fma.f32x2 -> FADD2 / FMUL2 / FFMA2 on sm_100): a complex value is one 64-bit register
// pair, so a complex add is ONE instruction and a complex multiply TWO (ptxas folds the half swaps, the per-half
// negation and the scalar broadcasts into operand modifiers).  They issue at half the scalar rate (tools/ubench.cu:
// 1.8 vs 3.4 warp-instructions/clk/SM), i.e. the same FMA-pipe cycles per flop, but half the issue slots.
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
__device__ __forceinline__ float2 cmulc(float2 a, float tx, float ty) { return cmul(a, make_float2(tx, ty)); }
// c + a * (tx, ty): two packed FMAs
__device__ __forceinline__ float2 cfmac(float2 c, float2 a, float tx, float ty) {
    u64 m, rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(m) : "l"(pk(tx, tx)), "l"(pk(a.x, a.y)), "l"(pk(c.x, c.y)));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(pk(ty, ty)), "l"(pk(-a.y, a.x)), "l"(m));
    return up(rd);
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// output stores (features / map): written once, never read again by this kernel
__device__ __forceinline__ void st_out(float* p, float v) {
#if CD_STCS
    __stcs(p, v);
#else
    *p = v;
#endif
}
__device__ __forceinline__ void st_out(float2* p, float2 v) {
#if CD_STCS
    __stcs(p, v);
#else
    *p = v;
#endif
}
__device__ __forceinline__ void st_out(uint4* p, uint4 v) {
#if CD_STCS
    __stcs(p, v);
#else
    *p = v;
#endif
}

// the stem's map is read back by the very next kernel: plain stores, so that its last-written end can still be in L2
// (CD_MAP_STCS=1 gives them the streaming hint too, for comparison)
#ifndef CD_MAP_STCS
#define CD_MAP_STCS 0
#endif
__device__ __forceinline__ void st_map(uint4* p, uint4 v) {
#if CD_MAP_STCS
    __stcs(p, v);
#else
    *p = v;
#endif
}

// ----------------------------------------------------------------------------------------------
// 16-point DFT in registers, radix 4 x 4
// ----------------------------------------------------------------------------------------------
constexpr float kC1 = 0.92387953251128674f;     // cos(pi / 8)
constexpr float kS1 = 0.38268343236508977f;     // sin(pi / 8)
constexpr float kH  = 0.70710678118654752f;

// y_k = sum_j x_j (-i)^(j k), in place.  Z3: x3 is known to be zero (and is only written).
template <bool Z3>
__device__ __forceinline__ void r4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 t0 = cadd(x0, x2), t1 = csub(x0, x2);
    if (Z3) {
        const float2 j = mul_mi(x1);
        x0 = cadd(t0, x1);
        x2 = csub(t0, x1);
        x1 = cadd(t1, j);
        x3 = csub(t1, j);
    } else {
        const float2 t2 = cadd(x1, x3), t3 = mul_mi(csub(x1, x3));
        x0 = cadd(t0, t2);
        x2 = csub(t0, t2);
        x1 = cadd(t1, t3);
        x3 = csub(t1, t3);
    }
}
__device__ __forceinline__ float2 w16_2(float2 a) { return scale2(kH, cadd(a, mul_mi(a))); }     // * W16^2 = h (1 - i)
__device__ __forceinline__ float2 w16_6(float2 a) { return scale2(-kH, csub(a, mul_mi(a))); }    // * W16^6 = -h (1 + i)

// X[t] = sum_a x[a] W16^(a t); input natural order, output X[t] at index F16(t) (the in-place digit reversal).
// ZPAD: x[13], x[14], x[15] are zero (the 512-point frame holds 400 samples: a = 13..15 is padding).
#define F16(t) (4 * ((t) & 3) + ((t) >> 2))
template <bool ZPAD>
__device__ __forceinline__ void fft16(float2 (&x)[16]) {
    r4<false>(x[0], x[4], x[8], x[12]);
    r4<ZPAD>(x[1], x[5], x[9], x[13]);
    r4<ZPAD>(x[2], x[6], x[10], x[14]);
    r4<ZPAD>(x[3], x[7], x[11], x[15]);
    x[5] = cmulc(x[5], kC1, -kS1);            // u[a0][t0] *= W16^(a0 t0), u at index a0 + 4 t0
    x[9] = w16_2(x[9]);
    x[13] = cmulc(x[13], kS1, -kC1);
    x[6] = w16_2(x[6]);
    x[10] = mul_mi(x[10]);
    x[14] = w16_6(x[14]);
    x[7] = cmulc(x[7], kS1, -kC1);
    x[11] = w16_6(x[11]);
    x[15] = cmulc(x[15], -kC1, kS1);
    r4<false>(x[0], x[1], x[2], x[3]);
    r4<false>(x[4], x[5], x[6], x[7]);
    r4<false>(x[8], x[9], x[10], x[11]);
    r4<false>(x[12], x[13], x[14], x[15]);
}

// Stage 1 of the 512-point FFT: fft16<true> followed by the twiddles W512^(m r) of lane m, read from the lane's row of
// the shared-memory table.  The table is stored in the order of the in-place result (entry 4 t0 + t1 = W^(m (t0 + 4 t1))
// multiplies x[4 t0 + t1] = X[t0 + 4 t1]), and the loads of a group are issued two butterflies ahead of their use.
__device__ __forceinline__ void lds_tw4(uint32_t addr, float2 (&t)[4]) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t[0].x), "=f"(t[0].y), "=f"(t[1].x), "=f"(t[1].y) : "r"(addr));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t[2].x), "=f"(t[2].y), "=f"(t[3].x), "=f"(t[3].y) : "r"(addr + 16));
}
__device__ __forceinline__ void fft16_twiddled(float2 (&x)[16], uint32_t tw_s) {
    float2 ta[4], tb[4];
    lds_tw4(tw_s, ta);                          // group 0: r = 0, 4, 8, 12
    lds_tw4(tw_s + 32, tb);                     // group 1: r = 1, 5, 9, 13
    r4<false>(x[0], x[4], x[8], x[12]);
    r4<true>(x[1], x[5], x[9], x[13]);
    r4<true>(x[2], x[6], x[10], x[14]);
    r4<true>(x[3], x[7], x[11], x[15]);
    x[5] = cmulc(x[5], kC1, -kS1);
    x[9] = w16_2(x[9]);
    x[13] = cmulc(x[13], kS1, -kC1);
    x[6] = w16_2(x[6]);
    x[10] = mul_mi(x[10]);
    x[14] = w16_6(x[14]);
    x[7] = cmulc(x[7], kS1, -kC1);
    x[11] = w16_6(x[11]);
    x[15] = cmulc(x[15], -kC1, kS1);
    r4<false>(x[0], x[1], x[2], x[3]);
    x[1] = cmul(x[1], ta[1]);                   // x[0] = X[0]: W^0 = 1
    x[2] = cmul(x[2], ta[2]);
    x[3] = cmul(x[3], ta[3]);
    lds_tw4(tw_s + 64, ta);                     // group 2
    r4<false>(x[4], x[5], x[6], x[7]);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[4 + i] = cmul(x[4 + i], tb[i]);
    lds_tw4(tw_s + 96, tb);                     // group 3
    r4<false>(x[8], x[9], x[10], x[11]);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[8 + i] = cmul(x[8 + i], ta[i]);
    r4<false>(x[12], x[13], x[14], x[15]);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[12 + i] = cmul(x[12 + i], tb[i]);
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
// Two block-wide reductions (both max or both sum) over kFeatWarps warps with one pair of barriers; every thread
// gets the results.  `red` needs >= 16 floats and is safe to reuse after the call returns.
template <bool IS_MAX>
__device__ __forceinline__ float2 block_reduce2(float2 v, float* red) {
    v.x = IS_MAX ? warp_max(v.x) : warp_sum(v.x);
    v.y = IS_MAX ? warp_max(v.y) : warp_sum(v.y);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[warp] = v.x;
        red[kFeatWarps + warp] = v.y;
    }
    __syncthreads();
    const int src = lane & 7;                  // every group of 8 lanes reduces the (<= 8) per-warp values
    float a = red[src < kFeatWarps ? src : 0], b = red[kFeatWarps + (src < kFeatWarps ? src : 0)];
    if (!IS_MAX && src >= kFeatWarps) a = b = 0.f;    // (a repeated value is harmless for a max)
#pragma unroll
    for (int o = 4; o; o >>= 1) {
        const float oa = __shfl_xor_sync(0xffffffffu, a, o), ob = __shfl_xor_sync(0xffffffffu, b, o);
        a = IS_MAX ? fmaxf(a, oa) : a + oa;
        b = IS_MAX ? fmaxf(b, ob) : b + ob;
    }
    __syncthreads();
    return make_float2(a, b);
}

// AmplitudeToDB('power'): 10 log10(max(x, 1e-10)) (torchaudio functional.py:390) via MUFU.LG2
__device__ __forceinline__ float power_to_db(float x) {
    float lg;                                                      // the argument is >= 1e-10: no denormal handling needed
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(fmaxf(x, 1e-10f)));
    return 3.01029995663981195f * lg;                              // 10 * log10(2) * log2(x)
}

// za = Z[k], zb = Z[512 - k] of z = x1 + i x2: {|Z + conj Z'|^2, |Z - conj Z'|^2} = the two frames' power at bin k
// (the window carries the factor 1/2).  S = za + zb = (sr, di), D = za - zb = (dr, si).  Symmetric in (za, zb).
__device__ __forceinline__ float2 pair_power(float2 za, float2 zb) {
    const float2 S = cadd(za, zb), D = csub(za, zb);
    u64 t, rd;
    asm("mul.f32x2 %0, %1, %1;" : "=l"(t) : "l"(pk(D.y, D.x)));                         // (si^2, dr^2)
    asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(rd) : "l"(pk(S.x, S.y)), "l"(t));          // (sr^2 + si^2, di^2 + dr^2)
    return up(rd);
}

// ----------------------------------------------------------------------------------------------
// per-lane constants of the FFT / mel pipeline (all lane-constant for the life of the kernel)
// ----------------------------------------------------------------------------------------------
struct LaneCtx {
#if CD_TW == 0
    float2 tw[6];            // W512^(m r), r = 1, 2, 3, 4, 8, 12 (m = lane); the other twiddles are products of these
#endif
    uint32_t tw_s;           // shared address of this lane's row of the twiddle table
    float2 mw[kMelSlots];    // mel taps of the lane's two segments (cd_common.cuh)
    uint32_t mel_s[kMelSlots]; // shared addresses of the power pairs the lane's mel slots read (unused slots: the zero slot)
    uint32_t xw_s;           // shared address this lane writes exchange row 0 to
    uint32_t xr_s;           // shared address this lane reads its 16 stage-2 inputs from (128 B, contiguous)
    uint32_t pw_s;           // shared address of the warp's power region
    uint32_t p0_s, pb_s;     // where this lane's four power pairs go: slot 0, and slot j >= 1 at pb_s + j * pstep
    int pstep;
    int partner;             // lane holding Z[512 - k] for this lane's Z[k]
};

__device__ __forceinline__ void lane_ctx_init(LaneCtx& L, const FeatTables* __restrict__ tables, float2* xbw, const float2* twl, int lane) {
#if CD_TW == 0
    {
        constexpr int rs[6] = {1, 2, 3, 4, 8, 12};
#pragma unroll
        for (int i = 0; i < 6; ++i) L.tw[i] = tables->twl[lane][rs[i]];
    }
#endif
    L.tw_s = smem_u32(twl + lane * kTwPitch);
#pragma unroll
    for (int s = 0; s < kMelSlots; ++s) L.mw[s] = tables->melt[s][lane];
    const int m = lane, c1 = m & 1;
    // stage-1 lane m = 2 b + c1 writes row r at r * 34 + c1 * 16 + (b ^ 8 c1): the reader of (r, c1) gets 16 contiguous
    // float2 (c1 = 1: its two halves swapped, which the final butterfly's constants absorb), writes are conflict-free
    L.xw_s = smem_u32(xbw + c1 * 16 + ((m >> 1) ^ (c1 * 8)));
    const int r = lane & 15, c = lane >> 4;
    L.xr_s = smem_u32(xbw + r * kXchRow + c * 16);
    L.pw_s = smem_u32(xbw + kXchElems);
    // power pairs {P_frame0, P_frame1} at padded bin index pidx(k) = k + 2 (k / 16):
    //   c = 0: bins k = r + 16 j (lane 0, j = 0: bin 64);  c = 1: bins 128 - r - 16 j  (r = 0, j = 0: bin 128, never read)
    int p0, pb, st;
    if (c == 0) {
        pb = r; st = 18; p0 = lane == 0 ? 72 : r;
    } else {
        pb = r ? 142 - r : 144; st = -18; p0 = pb;
    }
    L.p0_s = L.pw_s + 8u * (uint32_t)p0;
    L.pb_s = L.pw_s + 8u * (uint32_t)pb;
    L.pstep = 8 * st;
    L.partner = ((16 - r) & 15) + 16 * (1 - c);
    {
        // the lane's mel slots: bins b, b + 1, .. of its two segments at padded index k + 2 (k / 16); a slot beyond the segment
        // reads the region's zero slot instead of being predicated off
        const int lo = tables->melb[0][lane], hi = tables->melb[1][lane];
#pragma unroll
        for (int sl = 0; sl < kMelSlots; ++sl) {
            const int packed = sl < kMelSlotsLo ? lo : hi, k = sl < kMelSlotsLo ? sl : sl - kMelSlotsLo;
            const int bin = (packed & 0xff) + k;
            L.mel_s[sl] = L.pw_s + 8u * (uint32_t)(k < (packed >> 8) ? bin + 2 * (bin >> 4) : kPowZero);
        }
    }
#if CD_PIN_CTX
    // opaque to the compiler from here on: it would otherwise re-derive these lane constants from %tid inside the pair loop
    // (about 30 integer instructions per pair) instead of keeping six registers
    asm volatile("" : "+r"(L.xw_s), "+r"(L.xr_s), "+r"(L.tw_s), "+r"(L.p0_s), "+r"(L.pb_s), "+r"(L.pw_s));
    asm volatile("" : "+r"(L.pstep), "+r"(L.partner));
    asm volatile("" : "+r"(L.mel_s[0]), "+r"(L.mel_s[1]), "+r"(L.mel_s[2]), "+r"(L.mel_s[3]), "+r"(L.mel_s[4]), "+r"(L.mel_s[5]));
#endif
}

// One frame pair, first half: x[a] = z[32 a + lane], a = 0..12 (x[13..15] ignored) -> the pair's two power spectra
// (bins 4..127) in the warp's power region.  mel_pair() finishes the job; the caller issues the next pair's sample
// loads between the two (the FFT registers are dead by then).
__device__ __forceinline__ void fft_pair(float2 (&x)[16], const LaneCtx& L, int lane) {
#if CD_ABLATE == 3
    {   // keep the loads and the window multiplies alive, nothing else
        float2 acc = x[0];
#pragma unroll
        for (int a = 1; a < 13; ++a) acc = cadd(acc, x[a]);
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(L.p0_s), "f"(acc.x), "f"(acc.y) : "memory");
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(L.pb_s + L.pstep), "f"(acc.y), "f"(acc.x) : "memory");
        return;
    }
#endif
    // ---- stage 1: DFT16 over a, then twiddles W512^(m r) ---------------------------------------------
#if CD_TW == 2
    fft16_twiddled(x, L.tw_s);
#else
    fft16<true>(x);
    {
        const float2 w1 = L.tw[0], w2 = L.tw[1], w3 = L.tw[2], w4 = L.tw[3], w8 = L.tw[4], w12 = L.tw[5];
        x[F16(1)] = cmul(x[F16(1)], w1);
        x[F16(2)] = cmul(x[F16(2)], w2);
        x[F16(3)] = cmul(x[F16(3)], w3);
        x[F16(4)] = cmul(x[F16(4)], w4);
        x[F16(5)] = cmul(x[F16(5)], cmul(w4, w1));
        x[F16(6)] = cmul(x[F16(6)], cmul(w4, w2));
        x[F16(7)] = cmul(x[F16(7)], cmul(w4, w3));
        x[F16(8)] = cmul(x[F16(8)], w8);
        x[F16(9)] = cmul(x[F16(9)], cmul(w8, w1));
        x[F16(10)] = cmul(x[F16(10)], cmul(w8, w2));
        x[F16(11)] = cmul(x[F16(11)], cmul(w8, w3));
        x[F16(12)] = cmul(x[F16(12)], w12);
        x[F16(13)] = cmul(x[F16(13)], cmul(w12, w1));
        x[F16(14)] = cmul(x[F16(14)], cmul(w12, w2));
        x[F16(15)] = cmul(x[F16(15)], cmul(w12, w3));
    }
#endif
    // ---- the one exchange: lane (r, c) gathers Y'[2 b + c][r], b = 0..15 -------------------------------
#if CD_ABLATE != 4
#pragma unroll
    for (int r = 0; r < 16; ++r)
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(L.xw_s + r * (kXchRow * 8)), "f"(x[F16(r)].x), "f"(x[F16(r)].y) : "memory");
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(x[2 * j].x), "=f"(x[2 * j].y), "=f"(x[2 * j + 1].x), "=f"(x[2 * j + 1].y) : "r"(L.xr_s + j * 16) : "memory");
#endif
    // ---- stage 2: DFT16 over b.  c = 0: V(t) = E[t];  c = 1: V(t) = (-1)^t O[(t + 8) & 15] (sign flip of the
    //      lanes m = 3 mod 4 folded into their window + the swapped halves of the c = 1 rows) ----------------
    fft16<false>(x);
    // ---- radix 2 across the lane pair (r, 0) <-> (r, 1), pruned: c = 0 keeps Z[r + 16 i], c = 1 Z[384 + r + 16 i] ----
    float2 Z[8];
    {
        float2 R[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            R[i].x = __shfl_xor_sync(0xffffffffu, x[F16(8 + i)].x, 16);
            R[i].y = __shfl_xor_sync(0xffffffffu, x[F16(8 + i)].y, 16);
        }
        // T_i = (-1)^i W32^i = (-1)^i (c_i, -s_i);  K_i = i T_i = (-1)^i (s_i, c_i)
        constexpr float kc[8] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                                 0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f};
        constexpr float ks[8] = {0.f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f,
                                 0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f};
        if (lane < 16) {                                   // Z = V(i) + T_i R_i
            Z[0] = cadd(x[F16(0)], R[0]);
#pragma unroll
            for (int i = 1; i < 8; ++i) {
                const float sg = (i & 1) ? -1.f : 1.f;
                Z[i] = cfmac(x[F16(i)], R[i], sg * kc[i], -sg * ks[i]);
            }
        } else {                                           // Z = R_i + K_i V(i)
            Z[0] = csub(R[0], mul_mi(x[F16(0)]));
#pragma unroll
            for (int i = 1; i < 8; ++i) {
                const float sg = (i & 1) ? -1.f : 1.f;
                Z[i] = cfmac(R[i], x[F16(i)], sg * ks[i], sg * kc[i]);
            }
        }
    }
    // ---- unpack the two real spectra: Z[512 - k] lives in the partner lane ((16 - r) & 15, 1 - c); the two lanes
    //      swap their upper four values and each forms four bins' power pairs (tools/fft_lane_model.py) -------------
    {
        const bool r0 = (lane & 15) == 0;                  // residue 0: pairing shifted by one, bin 64 in lane 0's free slot
        float2 rq[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 a = Z[4 + q], b = Z[4 + ((q + 1) & 3)];
            rq[q].x = __shfl_sync(0xffffffffu, r0 ? b.x : a.x, L.partner);
            rq[q].y = __shfl_sync(0xffffffffu, r0 ? b.y : a.y, L.partner);
        }
        const float2 own0 = lane == 0 ? Z[4] : Z[0];
        const float2 p0 = pair_power(own0, rq[3]), p1 = pair_power(Z[1], rq[2]);
        const float2 p2 = pair_power(Z[2], rq[1]), p3 = pair_power(Z[3], rq[0]);
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(L.p0_s), "f"(p0.x), "f"(p0.y) : "memory");
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(L.pb_s + L.pstep), "f"(p1.x), "f"(p1.y) : "memory");
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(L.pb_s + 2 * L.pstep), "f"(p2.x), "f"(p2.y) : "memory");
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(L.pb_s + 3 * L.pstep), "f"(p3.x), "f"(p3.y) : "memory");
    }
}

// Second half: mel POWER of the two frames for bands `lane` (lo) and `lane + 32` (hi); .x = first frame, .y = second.
__device__ __forceinline__ void mel_pair(const LaneCtx& L, int lane, float2& lo, float2& hi) {
    __syncwarp();
#if CD_ABLATE == 2 || CD_ABLATE == 3
    lo = lds64(L.p0_s);
    hi = lds64(L.pb_s + L.pstep);
    return;
#endif
    // ---- banded mel contraction (cd_common.cuh): the lane reads the bins of its two segments once, for both
    //      frames, and hands the rising-side sums to the lane above ----------------------------------------
    lo = make_float2(0.f, 0.f);
    hi = lo;
    float2 ulo = lo, uhi = lo;
    {
#pragma unroll
        for (int s = 0; s < kMelSlotsLo; ++s) {
            const float2 pw = lds64(L.mel_s[s]);
            lo = fma2s(L.mw[s].x, pw, lo);
            ulo = fma2s(L.mw[s].y, pw, ulo);
        }
#pragma unroll
        for (int s = 0; s < kMelSlotsHi; ++s) {
            const float2 pw = lds64(L.mel_s[kMelSlotsLo + s]);
            hi = fma2s(L.mw[kMelSlotsLo + s].x, pw, hi);
            uhi = fma2s(L.mw[kMelSlotsLo + s].y, pw, uhi);
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
    // no trailing __syncwarp: the next pair's exchange writes touch another region, and its power stores come
    // after its own first __syncwarp, which every lane reaches only after these reads
}

// ----------------------------------------------------------------------------------------------
// sample access
// ----------------------------------------------------------------------------------------------
// The conditioned clip of AudioPreprocessor.process (:508-515) as a view of the caller's buffer: sample i of the
// 16 000 is src[start + i - padl] when 0 <= i - padl < len (centre trim: start > 0; zero pad: padl > 0), else 0.
template <class T>
struct ClipView {
    const T* src;
    int len, start, padl;
};
template <class T>
__device__ __forceinline__ float view_at(const ClipView<T>& v, int i) {
    const int j = i - v.padl;
    return (j >= 0 && j < v.len) ? (float)v.src[v.start + j] : 0.f;
}
// Sample i of the (possibly pre-emphasised) conditioned clip with torch.stft's reflect padding applied to the index.
// T = float (the reference's waveform type) or int16_t (PCM as stored in wav files: the power-of-two scale cancels in
// the peak normalisation, so both give bit-identical features).  PRE: y[i] = x[i] - coef x[i-1], y[0] = x[0]
// (apply_pre_emphasis, src/preprocessing.py:213-239; product and difference rounded separately like the tensor ops).
// WHOLE: the clip is exactly the 16 000 samples at v.src (no trim / pad mapping).
template <bool PRE, bool WHOLE, class T>
__device__ __forceinline__ float sample_reflect(const ClipView<T>& v, int i, float coef) {
    i = i < 0 ? -i : i;
    i = i >= kSeg ? 2 * (kSeg - 1) - i : i;
    const float x = WHOLE ? (float)v.src[i] : view_at(v, i);
    if (!PRE || i == 0) return x;
    return __fsub_rn(x, __fmul_rn(coef, WHOLE ? (float)v.src[i - 1] : view_at(v, i - 1)));
}

// window weights of lane m: w[a] = 0.5 hann[32 a + m] (0 beyond sample 399), times the lane's sign (lanes m = 3 mod 4
// carry -1: that rotates the second DFT16 of the odd lanes by 8, see fft_mel_pair) and the power-of-two pre-scale
__device__ __forceinline__ void lane_window(const FeatTables* __restrict__ tables, float scale, int lane, float (&w)[13]) {
    const float s = (lane & 3) == 3 ? -scale : scale;
#pragma unroll
    for (int a = 0; a < 12; ++a) w[a] = __ldg(&tables->win[32 * a + lane]) * s;
    w[12] = lane < 16 ? __ldg(&tables->win[384 + lane]) * s : 0.f;
}

// raw[q] = sample s0 + 32 q + lane of the pair's 576-sample span; frame 0 uses q = 0..12, frame 1 q = 5..17
__device__ __forceinline__ void form_inputs(const float (&raw)[18], const float (&w)[13], float2 (&x)[16]) {
#pragma unroll
    for (int a = 0; a < 13; ++a) x[a] = make_float2(w[a] * raw[a], w[a] * raw[a + 5]);
}
// the pair's mel power into the tile: frames 2 p, 2 p + 1 of bands lane and lane + 32 as two 8-byte stores
__device__ __forceinline__ float store_pair(float* db, int lane, int p, bool second, float2 lo, float2 hi) {
    float2* row_lo = reinterpret_cast<float2*>(db + lane * kTilePitch + 2 * p);
    float2* row_hi = reinterpret_cast<float2*>(db + (lane + 32) * kTilePitch + 2 * p);
    if (!second) lo.y = hi.y = 0.f;                                    // column 101 of the tile: never read
    *row_lo = lo;
    *row_hi = hi;
    return fmaxf(fmaxf(lo.x, hi.x), fmaxf(lo.y, hi.y));
}

// DCT-II of the floored dB tile for the frame pair (2 j, 2 j + 1), coefficients of one parity (PAR = 0: c = 0,2,..,12;
// PAR = 1: c = 1,3,..,11).  cos(pi (m + 1/2) c / 64) is even (c even) or odd (c odd) about m = 31.5, so
// each coefficient is a 32-term sum over d[m] +- d[63 - m], shared by all coefficients of that parity; both frames
// ride in one packed register pair (one 8-byte tile read per band, FFMA2).
// POWER: the tile holds mel power (PCEN variant); dB and its top_db floor are applied on the way in.
// Returns {sum, sum of squares} of the coefficients it wrote (the block's mean / std come from one reduction).
template <int PAR, bool POWER>
__device__ __forceinline__ float2 mfcc_par(const float* __restrict__ db, float* __restrict__ mf, int j, float floor_db) {
    constexpr int NC = PAR ? 6 : 7;
    float2 acc[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int m = 0; m < kMels / 2; ++m) {
        float2 lo = *reinterpret_cast<const float2*>(db + m * kTilePitch + 2 * j);
        float2 hi = *reinterpret_cast<const float2*>(db + (kMels - 1 - m) * kTilePitch + 2 * j);
        if (POWER) {
            lo = make_float2(fmaxf(power_to_db(lo.x), floor_db), fmaxf(power_to_db(lo.y), floor_db));
            hi = make_float2(fmaxf(power_to_db(hi.x), floor_db), fmaxf(power_to_db(hi.y), floor_db));
        }
        const float2 d = PAR ? csub(lo, hi) : cadd(lo, hi);
#pragma unroll
        for (int i = 0; i < NC; ++i) acc[i] = fma2s(c_dct[(PAR + 2 * i) * kMels + m], d, acc[i]);
    }
    const bool two = 2 * j + 1 < kFrames;                              // frame 101 does not exist
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        *reinterpret_cast<float2*>(mf + (PAR + 2 * i) * kTilePitch + 2 * j) = acc[i];
        s += acc[i].x;
        s2 = fmaf(acc[i].x, acc[i].x, s2);
        if (two) {
            s += acc[i].y;
            s2 = fmaf(acc[i].y, acc[i].y, s2);
        }
    }
    return make_float2(s, s2);
}

// Everything after the power tile is complete: peak scaling, dB, top_db floor, mel rows, DCT-13, mean / unbiased std
// normalisation, deltas; fp32 features to `out` and/or the stem's bf16 map to `mo` (either may be null).
// The tile holds un-normalised mel power, `top` is its block max, inv2 = 1 / peak^2 (normalize(), :199-212, applied to
// the power: the STFT is linear).
// PCEN = false: rows 0..63 are clamp((dB + 80) / 80) of the floored dB.
// PCEN = true : rows 0..63 are apply_pcen (src/preprocessing.py:305-340) scaled to [0, 1] by the clip's min / max
//               (:400-404); the MFCC rows still come from the dB.
// Ends with the block in sync except for the final stores.
// dd: also emit the delta-delta rows 90..102 = compute_deltas(compute_deltas(mfcc)) (:463-466).
template <bool PCEN, class SM>
__device__ __forceinline__ void feat_tail(SM& S, float top, float inv2, float* __restrict__ out, __nv_bfloat16* __restrict__ mo,
                                          int tid, bool dd = false) {
    const float floor_db = power_to_db(top * inv2) - 80.0f;           // top_db = 80 (monotone: the max of the dB)

    // ---- rows 0..63: P / peak^2 -> dB -> max(dB, floor) -> (dB + 80) / 80 -> clamp ------------------------
    // The CNN stem's operand rides along: the same values rounded to bf16 in the clip's blocked map (cd_common.cuh:
    // bmap_offset).  A 2-byte store to global memory costs a whole 32-byte sector AND makes the L2 fetch that sector
    // first (measured: +27 KB of DRAM reads per clip), so the map is STAGED in shared memory -- the FFT exchange
    // buffers are idle in the tail -- and copied out with 16-byte stores that cover full sectors.  Two stages (the
    // buffers hold 17 KB beside the MFCC scratch): block rows 0..15 (map rows 0..63: zero rows, mel rows 0..60) here,
    // the rest (mel rows 61..63, MFCC, deltas, zero rows) after the normalisation.  Every element of a staged block
    // row is written exactly once, zero padding included.  (PCEN keeps the scattered stores: its rows 0..63 are
    // known only at the very end.)
    constexpr int kStageOff = 5376;                                    // bytes into S.xb: behind the MFCC scratch
    constexpr int kStageRowsA = 16;                                    // block rows in stage A
    static_assert(kStageOff >= (int)(sizeof(float) * kMfcc * kTilePitch), "stage behind the MFCC scratch");
    static_assert(kStageOff + kStageRowsA * kBmapB * 32 <= (int)(sizeof(float2) * kXbElems * kFeatWarps), "stage A fits");
    static_assert(kStageOff + (NetShape<kFeatDD>::BA - kStageRowsA) * kBmapB * 32 <= (int)(sizeof(float2) * kXbElems * kFeatWarps), "stage B fits");
    __nv_bfloat16* stage = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(&S.xb[0][0]) + kStageOff);
    const bool staged = !PCEN && mo != nullptr;
    const __nv_bfloat16 bz = __float2bfloat16_rn(0.f);
    if (!PCEN) {
        // work item = (block row a of the map, frame pair j): four tile rows x two frames, so the blocked-map address
        // arithmetic is shared by the four rows and no lane idles (17 x 51 items over the CTA).  The top_db floor and the
        // 1e-10 clamp move to the power side of the logarithm:
        //   max(10 log10(max(Q, 1e-10)), floor) = kDb lg2(max(Q, qfloor)),  Q = P / peak^2,  qfloor = max(1e-10, top / peak^2 * 1e-8)
        // (monotone, so exact up to the rounding of the logarithm; Q itself is formed first, which keeps the features
        // bit-identical under power-of-two gains and between the int16 and fp32 entries), and (dB + 80) / 80 = dB / 80 + 1
        // is one FFMA.SAT.
        constexpr float kDb = 3.01029995663981195f;
        const float qfloor = fmaxf(1e-10f, (top * inv2) * 1e-8f);
        const bool al8 = out && (reinterpret_cast<uintptr_t>(out) & 7) == 0;
        for (int idx = tid; idx < (kMels / 4 + 1) * kPairs; idx += kFeatThreads) {
            const int a = idx / kPairs, j = idx - a * kPairs;
            const bool two = j < kPairs - 1;
            // blocked-map addresses of the item's two columns (cd_common.cuh bmap_offset, unrolled over the block's four
            // rows): element (4 a + rr, C) sits at row * 16 + (((rr >> 1) << 3) ^ swz) + (rr & 1) * 4 + (C & 3)
            const int C0 = 2 * j + 3, C1 = 2 * j + 4;
            const int row0 = a * kBmapB + (C0 >> 2), row1 = a * kBmapB + (C1 >> 2);
            __nv_bfloat16* const m0 = stage + row0 * 16 + (C0 & 3);
            __nv_bfloat16* const m1 = stage + row1 * 16 + (C1 & 3);
            const int z0 = ((row0 >> 2) & 1) << 3, z1 = ((row1 >> 2) & 1) << 3;
            const bool do_stage = staged && a < kStageRowsA;
            float* const o = out + (4 * a - 3) * kFrames + 2 * j;
            const float2* cellp = reinterpret_cast<const float2*>(S.db + (4 * a - 3) * kTilePitch + 2 * j);
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int r = 4 * a + rr - 3;
                float2 x01 = make_float2(0.f, 0.f);
                if (r >= 0 && r < kMels) {
                    float2* cell = const_cast<float2*>(cellp) + rr * (kTilePitch / 2);
                    const float2 p = *cell;
                    float lx, ly;
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lx) : "f"(fmaxf(p.x * inv2, qfloor)));
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(ly) : "f"(fmaxf(p.y * inv2, qfloor)));
                    const float2 v = make_float2(kDb * lx, kDb * ly);
                    *cell = v;
                    x01 = make_float2(__saturatef(fmaf(v.x, 0.0125f, 1.0f)), __saturatef(fmaf(v.y, 0.0125f, 1.0f)));
                    if (out) {
                        float* orow = o + rr * kFrames;
                        // row r of an 8-byte aligned output starts 8-byte aligned when r is even (101 floats per row)
                        if (al8 && !((rr + a * 4 - 3) & 1) && two) {
                            st_out(reinterpret_cast<float2*>(orow), x01);
                        } else {
                            st_out(orow, x01.x);
                            if (two) st_out(orow + 1, x01.y);
                        }
                    }
                }
                if (do_stage) {
                    const int e0 = (((rr >> 1) << 3) ^ z0) + (rr & 1) * 4, e1 = (((rr >> 1) << 3) ^ z1) + (rr & 1) * 4;
                    m0[e0] = __float2bfloat16_rn(x01.x);
                    m1[e1] = two ? __float2bfloat16_rn(x01.y) : bz;                                // C = 104: right padding
                    if (j == 0) {                                                                  // C = 0, 1, 2 sit left of C0 = 3
                        m0[e0 - 3] = bz; m0[e0 - 2] = bz; m0[e0 - 1] = bz;
                    }
                    if (j == kPairs - 1) {                                                         // C = 105, 106, 107 right of C1 = 104
                        m1[e1 + 1] = bz; m1[e1 + 2] = bz; m1[e1 + 3] = bz;
                    }
                }
            }
        }
    } else {
        // PCEN: the tile keeps the scaled mel POWER (the MFCC rows take their dB from it on the fly)
        for (int idx = tid; idx < kMels * kPairs; idx += kFeatThreads) {
            const int r = idx / kPairs, j = idx - r * kPairs;
            float2* cell = reinterpret_cast<float2*>(S.db + r * kTilePitch + 2 * j);
            float2 p = *cell;
            p.x *= inv2;
            p.y *= inv2;
            *cell = p;
        }
    }
    __syncthreads();
    if (staged) {                                                      // stage A -> global, full sectors
        const uint4* s16 = reinterpret_cast<const uint4*>(stage);
        uint4* g16 = reinterpret_cast<uint4*>(mo);
        for (int i = tid; i < kStageRowsA * kBmapB * 2; i += kFeatThreads) st_map(g16 + i, s16[i]);
    }

    // ---- MFCC = DCT-II(13 x 64) of the floored dB, then (x - mean) / (std + 1e-8) ---------------
    float* mf = reinterpret_cast<float*>(&S.xb[0][0]);             // [13][102], FFT scratch is idle now
    static_assert(sizeof(float) * kMfcc * kTilePitch <= sizeof(float2) * (kXchElems + kPowZero), "MFCC scratch fits under warp 0's zero slot");
    float2 part = make_float2(0.f, 0.f);
    {
        const int j = tid & 63, g = tid >> 6;                      // warps 0-1: even coefficients, warps 2-3: odd ones
        if (j < kPairs && g < 2) part = g ? mfcc_par<1, PCEN>(S.db, mf, j, floor_db) : mfcc_par<0, PCEN>(S.db, mf, j, floor_db);
    }
    // mean and unbiased variance from one reduction of {sum, sum of squares}: |mean| is a fraction of the spread
    // here (c0 dominates both), so the one-pass form loses nothing in fp32 (checked against float64 on the goldens)
    constexpr int kN = kMfcc * kFrames;                            // 1313
    const float2 tot = block_reduce2<false>(part, S.red);
    const float mean = tot.x * (1.0f / kN);
    const float var = fmaxf(tot.y - tot.x * mean, 0.f) * (1.0f / (kN - 1));
    const float inv_std = 1.0f / (sqrtf(var) + 1e-8f);
    // stage B holds map rows 64 .. 4 BA - 1 (block rows 16 ..); element (R, C) sits at bmap_offset(R, C) - kStageBase
    constexpr int kStageBase = kStageRowsA * kBmapB * 16;
    const int map_rows = dd ? 4 * NetShape<kFeatDD>::BA : 4 * kBmapA;  // 112 | 96
    {
        // thread = (coefficient, frame pair): normalised MFCC rows 64.., their deltas 77.. (compute_deltas, :342-356:
        // replicate padding, (x[t+1] - x[t-1]) / 2) and on request the deltas of the deltas 90..; 13 x 51 items, the pair's
        // four neighbours read once.  __fmul_rn: never contracted into an FMA, so every consumer sees identical values.
        const bool al8o = out && (reinterpret_cast<uintptr_t>(out) & 7) == 0;
        __nv_bfloat16* const m = mo ? (staged ? stage - kStageBase : mo) : nullptr;
        for (int idx = tid; idx < kMfcc * kPairs; idx += kFeatThreads) {
            const int cidx = idx / kPairs, j = idx - cidx * kPairs, t = 2 * j;
            const bool two = j < kPairs - 1;
            const float* row = mf + cidx * kTilePitch;
            auto nrm = [&](float v) { return __fmul_rn(v - mean, inv_std); };
            const float2 pr = *reinterpret_cast<const float2*>(row + t);
            const float xa = nrm(pr.x);                                    // frame t
            const float xb = two ? nrm(pr.y) : xa;                         // frame t + 1 (replicated past the last frame)
            const float xm = t ? nrm(row[t - 1]) : xa;                     // frame t - 1 (replicated before the first)
            const float xq = nrm(row[min(t + 2, kFrames - 1)]);            // frame t + 2
            const float da = (xb - xm) * 0.5f, db = (xq - xa) * 0.5f;
            float ea = 0.f, eb = 0.f;
            if (dd) {
                auto xn = [&](int u) { return nrm(row[min(max(u, 0), kFrames - 1)]); };
                auto d1 = [&](int u) { u = min(max(u, 0), kFrames - 1); return (xn(u + 1) - xn(u - 1)) * 0.5f; };
                ea = (d1(t + 1) - d1(t - 1)) * 0.5f;
                eb = (d1(t + 2) - d1(t)) * 0.5f;
            }
            auto put = [&](int frow, float v0, float v1) {                 // feature row frow, frames t and t + 1
                if (out) {
                    float* o = out + frow * kFrames + t;
                    if (al8o && !(frow & 1) && two) {
                        st_out(reinterpret_cast<float2*>(o), make_float2(v0, v1));
                    } else {
                        st_out(o, v0);
                        if (two) st_out(o + 1, v1);
                    }
                }
                if (m) {
                    m[bmap_offset(frow + 3, t + 3)] = __float2bfloat16_rn(v0);
                    if (two) m[bmap_offset(frow + 3, t + 4)] = __float2bfloat16_rn(v1);
                }
            };
            put(kMels + cidx, xa, xb);
            put(kMels + kMfcc + cidx, da, db);
            if (dd) put(kMels + 2 * kMfcc + cidx, ea, eb);
        }
    }
    if (staged) {
        // the rest of stage B, thread = frame column: mel rows 61..63 (map rows 64..66) and the zero rows below the features
        __nv_bfloat16* m = stage - kStageBase;
        for (int t = tid; t < kFrames; t += kFeatThreads) {
#pragma unroll
            for (int r = 4 * kStageRowsA - 3; r < kMels; ++r) {
                const float v = S.db[r * kTilePitch + t];
                m[bmap_offset(r + 3, t + 3)] = __float2bfloat16_rn(__saturatef(fmaf(v, 0.0125f, 1.0f)));   // as the rows pass
            }
            for (int R = (dd ? kFeatDD : kFeat) + 3; R < map_rows; ++R) m[bmap_offset(R, t + 3)] = bz;
        }
    }
    if (staged) {
        // left / right zero columns of stage B (C = 0, 1, 2 and 104..107: the last block column is all padding): one thread
        // per (row, side), the four right-hand cells as one 8-byte store, the three left-hand ones as 4 + 2 bytes
        for (int i = tid; i < 2 * (map_rows - 4 * kStageRowsA); i += kFeatThreads) {
            const int R = 4 * kStageRowsA + (i >> 1);
            if (i & 1) {
                *reinterpret_cast<uint2*>(stage + (bmap_offset(R, 104) - kStageBase)) = make_uint2(0u, 0u);
            } else {
                __nv_bfloat16* l = stage + (bmap_offset(R, 0) - kStageBase);
                *reinterpret_cast<uint32_t*>(l) = 0u;
                l[2] = bz;
            }
        }
        __syncthreads();
        const int n16 = (map_rows / 4 - kStageRowsA) * kBmapB * 2;    // 16-byte pieces of stage B
        const uint4* s16 = reinterpret_cast<const uint4*>(stage);
        uint4* g16 = reinterpret_cast<uint4*>(mo) + kStageRowsA * kBmapB * 2;
        for (int i = tid; i < n16; i += kFeatThreads) st_map(g16 + i, s16[i]);
    }

    if (PCEN) {
        // ---- PCEN rows 0..63 from the power tile: smoother = zero-padded mean of frames t-5 .. t+4 (avg_pool2d
        //      kernel (1, 10), padding (0, 5), cut to 101 columns); (P / (eps + M)^0.98 + 2)^0.5 - 2^0.5 ------------
        constexpr int kPer = (kMels * kFrames + kFeatThreads - 1) / kFeatThreads;      // 26 (8 warps)
        float val[kPer];
        float vmin = INFINITY, vmax = -INFINITY;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int i = tid + k * kFeatThreads;
            float v = 0.f;
            if (i < kMels * kFrames) {
                const int r = i / kFrames, t = i - r * kFrames;
                const float* row = S.db + r * kTilePitch;
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
        const float2 mm = block_reduce2<true>(make_float2(vmax, -vmin), S.red);
        vmax = mm.x;
        vmin = -mm.y;
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
__device__ __forceinline__ ClipView<T> clip_view(const T* audio, const int64_t* offs, const int32_t* lens,
                                                int64_t stride_or_hop, int mode, int64_t b) {
    ClipView<T> v;
    if ((mode & kModeMask) == kModeVar) {
        const int n = lens[b];
        v.src = audio + offs[b];
        v.len = n < kSeg ? n : kSeg;                                   // samples that survive pad_or_trim
        v.start = n > kSeg ? (n - kSeg) / 2 : 0;                       // trim offset (src/preprocessing.py:378)
        v.padl = n < kSeg ? (kSeg - n) / 2 : 0;                        // left zero pad (:383)
    } else {
        v.src = audio + b * stride_or_hop;
        v.len = kSeg;
        v.start = 0;
        v.padl = 0;
    }
    return v;
}

// The next pair's ten new samples (raw[8..17]); the loads are only ISSUED here.  Without pre-emphasis nothing
// consumes them until that pair starts (the peak tracking of these values happens there too, see feat_kernel): a
// consumer right behind the loads would stall the warp for the whole memory latency.
template <bool PRE, class T>
__device__ __forceinline__ void request_next_impl(const T* __restrict__ ps, float (&raw)[18], float& amax, float coef) {
#pragma unroll
    for (int q = 8; q < 18; ++q) {
        const float xq = (float)ps[32 * q];
        if (PRE) {
            amax = fmaxf(amax, fabsf(xq));
            raw[q] = __fsub_rn(xq, __fmul_rn(coef, (float)ps[32 * q - 1]));
        } else {
            raw[q] = xq;
        }
    }
}

// Frame pairs of a clip are dealt to the warps as CONTIGUOUS runs (8 warps: 6, 6, 7, 7, 7, 6, 6, 6) so that consecutive
// pairs of a warp share samples in registers; the longer runs go to middle warps (the first and last warp hold the
// slower reflect-padded pairs 0, 49, 50).
__device__ __forceinline__ void warp_pair_range(int warp, int& p0, int& p1) {
    constexpr int base = kPairs / kFeatWarps, rem = kPairs % kFeatWarps, e0 = (kFeatWarps - rem) / 2;
    const int extra_before = min(max(warp - e0, 0), rem);
    p0 = warp * base + extra_before;
    p1 = p0 + base + ((warp >= e0 && warp < e0 + rem) ? 1 : 0);
}

// one-time per CTA: the per-lane twiddles that live in shared memory (CD_TW = 1)
template <class SM>
__device__ __forceinline__ void stage_tables(SM& S, const FeatTables* __restrict__ tables, int tid) {
    // entry 4 t0 + t1 of a lane's row = W^(m (t0 + 4 t1)): the order fft16_twiddled consumes them in
    for (int i = tid; i < 32 * 16; i += kFeatThreads) S.twl[i >> 4][i & 15] = tables->twl[i >> 4][((i & 15) >> 2) + 4 * (i & 3)];
}

// PRE / PCEN: the use_pre_emphasis / use_pcen variants of the 90-feature path (the part that keeps
// the CNN input at 90 rows); <false, false> is the training configuration (src/train.py:264-287) and the benchmark.
template <class T, bool PRE, bool PCEN>
__global__ void __launch_bounds__(kFeatThreads, kFeatCtasPerSm)
feat_kernel(const T* __restrict__ audio, const int64_t* __restrict__ offs, const int32_t* __restrict__ lens,
            int64_t stride_or_hop, int mode, int64_t B, float* __restrict__ feat,
            __nv_bfloat16* __restrict__ map_out, const FeatTables* __restrict__ tables, float coef) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FeatSmem& S = *reinterpret_cast<FeatSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    LaneCtx L;
    lane_ctx_init(L, tables, S.xb[warp], &S.twl[0][0], lane);
    stage_tables(S, tables, tid);
    if (lane == 0) S.xb[warp][kXchElems + kPowZero] = make_float2(0.f, 0.f);
    __syncthreads();
    int wp0, wp1;
    warp_pair_range(warp, wp0, wp1);
#if CD_PIN_CTX
    asm volatile("" : "+r"(wp0), "+r"(wp1));       // (kept in registers: otherwise re-derived from %tid in every pair)
#endif
    float w[13];                                                       // the lane's window weights, loaded once
    lane_window(tables, 1.0f, lane, w);
    const bool var = (mode & kModeMask) == kModeVar;
    const bool normalize = !(mode & kModeNoNormalize);

    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        if (lane == 0) S.xb[warp][kXchElems + kPowZero] = make_float2(0.f, 0.f);   // (the tail's map stage may have used it)
        const ClipView<T> v = clip_view(audio, offs, lens, stride_or_hop, mode, b);
        const bool fast = v.len == kSeg && v.start == 0;               // a whole 16 000-sample clip: direct loads
        // pull the clip after this one into L2 meanwhile (CD_L2PF says when)
        auto prefetch_next = [&](int part, int parts) {
            const int64_t bn = b + gridDim.x;
            if (bn < B) {
                const ClipView<T> vn = clip_view(audio, offs, lens, stride_or_hop, mode, bn);
                const T* pn = vn.src + vn.start;
                if (vn.len == kSeg && (reinterpret_cast<uintptr_t>(pn) & 15) == 0) {
                    const uint32_t bytes = (uint32_t)(kSeg * sizeof(T)) / parts;          // 64 000 / 4 and 32 000 / 4: multiples of 16
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(pn) + (size_t)part * bytes), "r"(bytes) : "memory");
                }
            }
        };
#if CD_L2PF == 1
        if (tid == 0) prefetch_next(0, 1);
#endif
        // AudioPreprocessor.process normalises over the WHOLE clip before it trims or pads: clips that do not take the
        // direct-load path get their peak from a pass of their own
        float peak_pre = 0.f;
        if (!fast && normalize) {
            const int n = var ? lens[b] : kSeg;
            const T* p = var ? audio + offs[b] : v.src;
            for (int i = tid; i < n; i += kFeatThreads) peak_pre = fmaxf(peak_pre, fabsf((float)p[i]));
        }

        float pre = 1.0f, top, peak;
        for (int pass = 0;; ++pass) {
            if (pass) lane_window(tables, pre, lane, w);               // (rare) pre-scaled second pass
            float pmax = 0.f, amax = peak_pre;
            float raw[18];
            // Reflected sample s (torch.stft's reflect padding) of a whole clip and, with pre-emphasis, the peak of the RAW
            // samples it touches (reflected samples are samples: they count for the peak).
            auto reflect_fast = [&](int s) {
                if (PRE) {
                    int i = s < 0 ? -s : s;
                    i = i >= kSeg ? 2 * (kSeg - 1) - i : i;
                    amax = fmaxf(amax, fabsf((float)v.src[i]));
                }
                return sample_reflect<PRE, true>(v, s, coef);
            };
            // ---- prologue: the 18 samples per lane of the run's first pair.  Every later pair takes its first eight from
            //      the pair before it (registers) and requests ten new ones a pair ahead, so the loop below has ONE shape
            //      for edge and interior pairs alike. -----------------------------------------------------------------
            {
                const int s0 = 2 * kHop * wp0 - kWin / 2 + lane;
                if (!fast) {                                           // trimmed / padded clip: everything through the index mapping
#pragma unroll
                    for (int q = 0; q < 18; ++q) raw[q] = sample_reflect<PRE, false>(v, s0 + 32 * q, coef);
                } else if (wp0 == 0 || wp0 >= 49) {                    // pair 0 starts in the left reflect padding (a run that
                                                                       // started at pair 49 / 50 would reach into the right one)
#pragma unroll
                    for (int q = 0; q < 18; ++q) raw[q] = reflect_fast(s0 + 32 * q);
                } else {
                    const T* ps = v.src + s0;
#pragma unroll
                    for (int q = 0; q < 18; ++q) {
                        const float xq = (float)ps[32 * q];
                        if (PRE) amax = fmaxf(amax, fabsf(xq));
                        raw[q] = PRE ? __fsub_rn(xq, __fmul_rn(coef, (float)ps[32 * q - 1])) : xq;
                    }
                }
                if (fast && !PRE) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) amax = fmaxf(amax, fabsf(raw[q]));      // the loop counts raw[8..17]
                }
            }
#pragma unroll kPairUnroll
            for (int p = wp0; p < wp1; ++p) {
#if CD_L2PF == 3
                if (lane == 0 && pass == 0 && p == (wp0 + wp1) / 2) prefetch_next(warp, kFeatWarps);
#endif
                const bool second = (2 * p + 1) < kFrames;             // pair 50 holds frame 100 only: its second frame is computed
                                                                       // from reflected samples like any other and dropped by store_pair
                const int n0 = 2 * kHop * (p + 1) - kWin / 2 + lane;   // the NEXT pair's sample of window tap 0
                float2 x[16], lo, hi;
                if (fast && !PRE) {                                    // requested during the previous pair: they arrive about now
#pragma unroll
                    for (int q = 8; q < 18; ++q) amax = fmaxf(amax, fabsf(raw[q]));
                }
                form_inputs(raw, w, x);
                // ---- the next pair's samples: 8 carried in registers, 10 new ones requested before the FFT and consumed
                //      after the mel contraction ---------------------------------------------------------------------
#pragma unroll
                for (int q = 0; q < 8; ++q) raw[q] = raw[q + 10];
                if (p + 1 < wp1) {
                    if (!fast) {
#pragma unroll
                        for (int q = 8; q < 18; ++q) raw[q] = sample_reflect<PRE, false>(v, n0 + 32 * q, coef);
                    } else if (p + 1 >= 49) {                          // pairs 49 and 50 reach into the right reflect padding
#pragma unroll
                        for (int q = 8; q < 18; ++q) raw[q] = reflect_fast(n0 + 32 * q);
                    } else {
                        request_next_impl<PRE>(v.src + n0, raw, amax, coef);
                    }
                }
                fft_pair(x, L, lane);
                mel_pair(L, lane, lo, hi);
                pmax = fmaxf(pmax, store_pair(S.db, lane, p, second, lo, hi));
            }
            const float2 mx = block_reduce2<true>(make_float2(pmax, amax), S.red);   // barrier: the power tile is complete
            top = mx.x;
            peak = mx.y;
            // normalize() (:199-212) is x / max|x|, skipped for silence.  A peak far from 1 would push the power out of
            // fp32's range, so such clips (|log2 peak| > 24) are redone with the window pre-scaled by 2^-e -- exact.
            if (pass == 0 && normalize && peak > 0.f && (peak < 5.9604645e-8f || peak > 16777216.f)) {
                pre = ldexpf(1.0f, min(max(-ilogbf(peak), -126), 126));
                continue;
            }
            break;
        }
        if (pre != 1.0f) lane_window(tables, 1.0f, lane, w);           // back to the plain window for the next clip
        float inv2 = 1.0f;
        if (normalize && peak > 0.f) {
            const float inv = 1.0f / (peak * pre);
            inv2 = inv * inv;
        }
#if CD_L2PF == 2
        if (tid == 0) prefetch_next(0, 1);
#endif
        const bool dd = (mode & kModeDeltaDelta) != 0;
#if CD_ABLATE == 1
        if (tid == 0 && feat) feat[b * (int64_t)kFeatElem] = top + inv2;
        if (false)
#endif
        feat_tail<PCEN>(S, top, inv2, feat ? feat + b * (int64_t)(dd ? kFeatElemDD : kFeatElem) : nullptr,
                        map_out ? map_out + b * (int64_t)(dd ? NetShape<kFeatDD>::MAP_ELEMS : kMapElems) : nullptr, tid, dd);
        __syncthreads();                                               // xb / db free again
    }
}

// ----------------------------------------------------------------------------------------------
// streaming with shared STFT frames (RealtimePreprocessor.add_audio, src/preprocessing.py:582-612)
// ----------------------------------------------------------------------------------------------
// With hop = 160 h, window j's frame t is the stream's global frame j h + t (centred at sample
// 160 (j h + t)); only frames 0, 1, 99, 100 of a window see its reflect padding.  So the mel POWER of
// every global frame is computed once (stream_mel_kernel), each window takes frames 2..97 from there
// and recomputes just the pairs (0,1), (98,99), (100,-) from its own samples; the peak normalisation is the
// scalar 1 / peak^2 on the power (feat_tail).  The window peak is the max of 100 per-hop block maxima.
struct MelSmem {
    float2 xb[kFeatWarps][kXbElems];
    float2 twl[32][kTwPitch];
};
using StreamSmem = FeatSmem;

// bmax[i] = max |x[160 i .. 160 i + 159]|: one warp per FOUR consecutive blocks, all twenty loads of a lane issued before the
// first is used (the kernel is pure latency: one block per warp kept 5 loads in flight and ran at a fifth of the HBM rate)
constexpr int kBmaxPerWarp = 4;
__global__ void __launch_bounds__(256)
block_max_kernel(const float* __restrict__ audio, int64_t n_blocks, float* __restrict__ bmax) {
    const int lane = threadIdx.x & 31;
    const int64_t i0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * kBmaxPerWarp;
    if (i0 >= n_blocks) return;
    const float* p = audio + i0 * kHop;
    const int64_t n = min((int64_t)kBmaxPerWarp, n_blocks - i0) * kHop;            // samples this warp covers
    float v[kBmaxPerWarp * kHop / 32];
#pragma unroll
    for (int k = 0; k < kBmaxPerWarp * kHop / 32; ++k) v[k] = 32 * k + lane < n ? fabsf(p[32 * k + lane]) : 0.f;
#pragma unroll
    for (int b = 0; b < kBmaxPerWarp; ++b) {
        float m = 0.f;
#pragma unroll
        for (int k = 0; k < kHop / 32; ++k) m = fmaxf(m, v[b * (kHop / 32) + k]);
        m = warp_max(m);
        if (lane == 0 && i0 + b < n_blocks) bmax[i0 + b] = m;
    }
}

// melp[f][64] = un-normalised mel power of global frame f, f in [2, F)
template <bool PRE>
__global__ void __launch_bounds__(kFeatThreads, kFeatCtasPerSm)
stream_mel_kernel(const float* __restrict__ audio, int64_t F, float* __restrict__ melp, const FeatTables* __restrict__ tables,
                  float coef) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MelSmem& S = *reinterpret_cast<MelSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    LaneCtx L;
    lane_ctx_init(L, tables, S.xb[warp], &S.twl[0][0], lane);
    stage_tables(S, tables, tid);
    if (lane == 0) S.xb[warp][kXchElems + kPowZero] = make_float2(0.f, 0.f);
    float w[13];
    lane_window(tables, 1.0f, lane, w);
    __syncthreads();
    const int64_t n_pairs = (F - 2 + 1) / 2;
    for (int64_t q = (int64_t)blockIdx.x * kFeatWarps + warp; q < n_pairs; q += (int64_t)gridDim.x * kFeatWarps) {
        int64_t f0 = 2 + 2 * q;
        if (f0 + 1 >= F) f0 = F - 2;                  // odd tail: redo (F-2, F-1); both writers store identical values
        const float* ps = audio + f0 * kHop - kWin / 2 + lane;       // frames >= 2 start >= 120 samples into the stream
        float raw[18];
#pragma unroll
        for (int k = 0; k < 18; ++k) {
            const float xq = ps[32 * k];
            raw[k] = PRE ? __fsub_rn(xq, __fmul_rn(coef, ps[32 * k - 1])) : xq;
        }
        float2 x[16], lo, hi;
        form_inputs(raw, w, x);
        fft_pair(x, L, lane);
        mel_pair(L, lane, lo, hi);
        float* o = melp + f0 * kMels + lane;
        o[0] = lo.x;
        o[32] = hi.x;
        o[kMels] = lo.y;
        o[kMels + 32] = hi.y;
    }
}

template <bool PRE, bool PCEN>
__global__ void __launch_bounds__(kFeatThreads, kFeatCtasPerSm)
stream_window_kernel(const float* __restrict__ audio, int64_t hop, int64_t n_windows, const float* __restrict__ melp,
                     const float* __restrict__ bmax, float* __restrict__ feat, __nv_bfloat16* __restrict__ map_out,
                     const FeatTables* __restrict__ tables, float coef, int dd) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StreamSmem& S = *reinterpret_cast<StreamSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    LaneCtx L;
    lane_ctx_init(L, tables, S.xb[warp], &S.twl[0][0], lane);
    stage_tables(S, tables, tid);
    if (lane == 0) S.xb[warp][kXchElems + kPowZero] = make_float2(0.f, 0.f);
    __syncthreads();
    const int64_t h = hop / kHop;
    float w[13];                                                       // the lane's window weights, loaded once
    lane_window(tables, 1.0f, lane, w);
    for (int64_t j = blockIdx.x; j < n_windows; j += gridDim.x) {
        if (lane == 0) S.xb[warp][kXchElems + kPowZero] = make_float2(0.f, 0.f);   // (the tail's map stage may have used it)
        const float peak = tid < kSeg / kHop ? bmax[j * h + tid] : 0.f;
        float wmax = 0.f;
        {
            // frames 2..97 come from the shared global frames j h + t.  Item = (mel band, frame pair): two coalesced loads
            // (consecutive lanes = consecutive bands of one frame) and one conflict-free 8-byte store into the band's tile
            // row -- by ALL warps, requested before the edge pairs below so the L2 latency hides under their FFTs.
            const float* src = melp + (j * h + 2) * kMels;
#pragma unroll 4
            for (int i = tid; i < 48 * kMels; i += kFeatThreads) {
                const int u = i >> 6, m = i & 63;
                const float d0 = src[(2 * u) * kMels + m], d1 = src[(2 * u + 1) * kMels + m];
                *reinterpret_cast<float2*>(S.db + m * kTilePitch + 2 * u + 2) = make_float2(d0, d1);
                wmax = fmaxf(wmax, fmaxf(d0, d1));
            }
        }
        if (warp < 3) {
            // frames that touch this window's reflect padding: pairs (0,1), (98,99), (100,-)
            const int p = warp == 0 ? 0 : 48 + warp;
            const bool second = (2 * p + 1) < kFrames;
            ClipView<float> v;
            v.src = audio + j * hop;
            v.len = kSeg;
            v.start = 0;
            v.padl = 0;
            const int s0 = 2 * kHop * p - kWin / 2 + lane;
            float raw[18];
#pragma unroll
            for (int q = 0; q < 18; ++q) raw[q] = sample_reflect<PRE, true>(v, s0 + 32 * q, coef);
            float2 x[16], lo, hi;
            form_inputs(raw, w, x);
            fft_pair(x, L, lane);
            mel_pair(L, lane, lo, hi);
            wmax = fmaxf(wmax, store_pair(S.db, lane, p, second, lo, hi));
        }
        const float2 mx = block_reduce2<true>(make_float2(wmax, peak), S.red);
        const float inv = mx.y > 0.f ? 1.0f / mx.y : 1.0f;           // normalize(): skipped for silence
        feat_tail<PCEN>(S, mx.x, inv * inv, feat ? feat + j * (int64_t)(dd ? kFeatElemDD : kFeatElem) : nullptr,
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
    {
        for (int m = 0; m < 32; ++m)
            for (int r = 0; r < 16; ++r) {
                const double a = -two_pi * (double)((m * r) % 512) / 512.0;
                t->twl[m][r] = make_float2((float)cos(a), (float)sin(a));
            }
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
    big(stream_mel_kernel<false>, sizeof(MelSmem));
    big(stream_mel_kernel<true>, sizeof(MelSmem));
    big(stream_window_kernel<false, false>, sizeof(StreamSmem));
    big(stream_window_kernel<true, false>, sizeof(StreamSmem));
    big(stream_window_kernel<false, true>, sizeof(StreamSmem));
    big(stream_window_kernel<true, true>, sizeof(StreamSmem));
    delete t;
    if (e != cudaSuccess) return set_err(h, CD_E_CUDA, "feature tables upload", e);
    return CD_OK;
}

// the four (pre-emphasis, PCEN) variants of a kernel template, picked by the call's options
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
    const int64_t slots = kFeatCtasPerSm * (int64_t)h->sm_count;     // resident CTAs
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
        block_max_kernel<<<(unsigned)((n_blocks + 8 * kBmaxPerWarp - 1) / (8 * kBmaxPerWarp)), 256, 0, s>>>(d_audio, n_blocks, bmax);
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
    int grid = (int)(B < slots ? B : slots);
    size_t smem = sizeof(FeatSmem);
#ifdef CD_FEAT_OCC_EXPERIMENT      // occupancy experiment (make exp): CD_OCC=<CTAs per SM> pads the dynamic shared memory
    if (const char* e = getenv("CD_OCC")) {
        const int occ = atoi(e);
        if (occ >= 1 && occ < kFeatCtasPerSm) {
            smem = (227 * 1024) / occ - 1024 - 512;
            grid = (int)(B < (int64_t)occ * h->sm_count ? B : (int64_t)occ * h->sm_count);
            cudaFuncSetAttribute(feat_kernel<float, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
    }
#endif
#define CD_CALL(P, Q)                                                                                          \
    feat_kernel<float, P, Q><<<grid, kFeatThreads, smem, s>>>(d_audio, d_offsets, d_lengths, stride_or_hop, \
                                                                          mode, B, d_feat, d_map, h->d_tables, coef)
    CD_FEAT_VARIANT(pre, pcen, CD_CALL);
#undef CD_CALL
    h->launches += 1;
    CD_CUDA(h, cudaGetLastError());
    return CD_OK;
}

// int16 PCM clips (fixed length): same kernel
int launch_features_i16(cd_handle* h, const int16_t* d_audio, int64_t stride, int mode, float coef, int64_t B,
                        float* d_feat, __nv_bfloat16* d_map, cudaStream_t s) {
    if (B == 0) return CD_OK;
    const int64_t slots = kFeatCtasPerSm * (int64_t)h->sm_count;
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
