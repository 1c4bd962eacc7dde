// CoughDetectorResidual forward on CUDA cores, fp32 (CD_PRECISION_FP32).
//
// Replaces src/model.py:253-259 (forward) with BatchNorm folded into the convolutions at load
// time (eval mode, Dropout identity).  Direct convolution, one CTA per clip, the clip's input
// activation staged zero-padded in shared memory.  This is the full-precision GPU path: its
// logits match the reference's fp32 forward to ~1e-5 and it is the on-device yardstick for the
// bf16 tcgen05 path (cnn_tc.cu).
//
//   stem   : 7x7/s2 conv (1->32) + bias + ReLU + 2x2 max-pool fused      (90,101) -> (32,22,25)
//   block a: 3x3/s2 conv + ReLU   and   1x1/s2 skip conv (shares the staged input)
//   block b: 3x3/s1 conv + skip add + ReLU   [+ global average pool + Linear(128->2) for block 1]
// Activations are NCHW fp32 in a workspace; layouts of the folded weights: [tap][cin][cout].
#include "cd_common.cuh"

namespace cd {

constexpr int kCnnThreads = 256;

// ---------------------------------------------------------------------------------------------
// stem: 7x7 stride 2 pad 3, 1 -> 32, ReLU, MaxPool2d(2) (floor mode: conv row 44 / col 50 unused)
// ---------------------------------------------------------------------------------------------
constexpr int kStemPitch = 106;    // input cols -3 .. 101 (+1)
constexpr int kPoolW = 25;
template <int NF>
struct StemF32 {
    static constexpr int POOLH = NetShape<NF>::H0;
    static constexpr int ROWS = (4 * POOLH + 5 + 1) & ~1; // rows 4 py + ky + 2 (second pool row) <= 4 POOLH + 4; even, so the
                                                          // weight table after the tile stays 16-byte aligned (94 | 110)
    static constexpr size_t SMEM = (size_t)(ROWS * kStemPitch + 49 * 32) * 4;
};

template <int NF>
__global__ void __launch_bounds__(kCnnThreads)
stem_kernel(const float* __restrict__ feat, const float* __restrict__ w /*[49][32]*/, const float* __restrict__ bias,
            float* __restrict__ x0 /*[B][32][H0][25]*/) {
    constexpr int kStemRows = StemF32<NF>::ROWS, kPoolH = StemF32<NF>::POOLH;
    extern __shared__ __align__(16) float sm[];
    float* tile = sm;                                   // [ROWS][106]
    float* ws = sm + kStemRows * kStemPitch;            // [49][32]
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.x;
    const float* f = feat + b * (int64_t)NetShape<NF>::FEAT_ELEMS;
    for (int i = tid; i < kStemRows * kStemPitch; i += kCnnThreads) {
        const int r = i / kStemPitch - 3, c = i % kStemPitch - 3;
        tile[i] = (r >= 0 && r < NF && c >= 0 && c < kFrames) ? f[r * kFrames + c] : 0.f;
    }
    for (int i = tid; i < 49 * 32; i += kCnnThreads) ws[i] = w[i];
    __syncthreads();

    // item = (pooled position, group of 8 output channels); channel group fastest across lanes
    for (int item = tid; item < kPoolH * kPoolW * 4; item += kCnnThreads) {
        const int cg = item & 3, pos = item >> 2;
        const int py = pos / kPoolW, px = pos - py * kPoolW;
        float acc[4][8];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[q][k] = 0.f;
        const float* base = tile + (4 * py) * kStemPitch + 4 * px;     // input (4py-3+ky, 4px-3+kx) at +ky,+kx
#pragma unroll 1
        for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const float* p = base + ky * kStemPitch + kx;
                const float v00 = p[0], v01 = p[2], v10 = p[2 * kStemPitch], v11 = p[2 * kStemPitch + 2];
                const float4 wa = *reinterpret_cast<const float4*>(ws + (ky * 7 + kx) * 32 + cg * 8);
                const float4 wb = *reinterpret_cast<const float4*>(ws + (ky * 7 + kx) * 32 + cg * 8 + 4);
                const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    acc[0][k] = fmaf(v00, wv[k], acc[0][k]);
                    acc[1][k] = fmaf(v01, wv[k], acc[1][k]);
                    acc[2][k] = fmaf(v10, wv[k], acc[2][k]);
                    acc[3][k] = fmaf(v11, wv[k], acc[3][k]);
                }
            }
        }
        float* o = x0 + (b * 32 + cg * 8) * (int64_t)(kPoolH * kPoolW) + pos;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float m = fmaxf(fmaxf(acc[0][k], acc[1][k]), fmaxf(acc[2][k], acc[3][k]));
            o[k * (kPoolH * kPoolW)] = fmaxf(m + bias[cg * 8 + k], 0.f);   // relu(max) == max(relu)
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 3x3 convolution (pad 1), optional fused 1x1/s2 skip conv, residual add, GAP + FC head
// ---------------------------------------------------------------------------------------------
template <int CIN, int COUT, int HIN, int WIN, int STRIDE, bool SKIP, bool RESID, bool HEAD>
__global__ void __launch_bounds__(kCnnThreads)
conv3x3_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
               const float* __restrict__ wsk, const float* __restrict__ bsk, const float* __restrict__ idt_in,
               float* __restrict__ y, float* __restrict__ idt_out, const float* __restrict__ fcw,
               const float* __restrict__ fcb, float* __restrict__ logits) {
    constexpr int HOUT = (HIN + 2 - 3) / STRIDE + 1, WOUT = (WIN + 2 - 3) / STRIDE + 1;
    constexpr int NPOS = HOUT * WOUT, NPG = (NPOS + 1) / 2, CG = COUT / 8;
    constexpr int HP = HIN + 2, WP = WIN + 2;
    extern __shared__ __align__(16) float sm[];
    float* tile = sm;                                   // [CIN][HP][WP], zero padded
    float* outs = sm + CIN * HP * WP;                   // HEAD: [COUT][NPOS]
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.x;
    const float* xin = x + b * (int64_t)(CIN * HIN * WIN);
    for (int i = tid; i < CIN * HP * WP; i += kCnnThreads) {
        const int ci = i / (HP * WP), r = (i / WP) % HP - 1, c = i % WP - 1;
        tile[i] = (r >= 0 && r < HIN && c >= 0 && c < WIN) ? xin[(ci * HIN + r) * WIN + c] : 0.f;
    }
    __syncthreads();

    for (int item = tid; item < NPG * CG; item += kCnnThreads) {
        const int cg = item % CG, pg = item / CG;
        const int p0 = 2 * pg, p1 = (2 * pg + 1 < NPOS) ? 2 * pg + 1 : p0;
        const int oy0 = p0 / WOUT, ox0 = p0 - oy0 * WOUT, oy1 = p1 / WOUT, ox1 = p1 - oy1 * WOUT;
        const float* t0 = tile + (oy0 * STRIDE) * WP + ox0 * STRIDE;
        const float* t1 = tile + (oy1 * STRIDE) * WP + ox1 * STRIDE;
        float a0[8], a1[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a0[k] = a1[k] = 0.f;
#pragma unroll 1
        for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int off = ci * HP * WP + (tap / 3) * WP + (tap % 3);
                const float v0 = t0[off], v1 = t1[off];
                const float* wp = w + ((tap * CIN + ci) * COUT + cg * 8);
                const float4 wa = __ldg(reinterpret_cast<const float4*>(wp));
                const float4 wb = __ldg(reinterpret_cast<const float4*>(wp + 4));
                const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    a0[k] = fmaf(v0, wv[k], a0[k]);
                    a1[k] = fmaf(v1, wv[k], a1[k]);
                }
            }
        }
        float s0[8], s1[8];
        if (SKIP) {     // 1x1 stride-2 projection reads the centre tap of the same staged input
#pragma unroll
            for (int k = 0; k < 8; ++k) s0[k] = s1[k] = 0.f;
#pragma unroll 4
            for (int ci = 0; ci < CIN; ++ci) {
                const int off = ci * HP * WP + WP + 1;
                const float v0 = t0[off], v1 = t1[off];
                const float* wp = wsk + (ci * COUT + cg * 8);
                const float4 wa = __ldg(reinterpret_cast<const float4*>(wp));
                const float4 wb = __ldg(reinterpret_cast<const float4*>(wp + 4));
                const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    s0[k] = fmaf(v0, wv[k], s0[k]);
                    s1[k] = fmaf(v1, wv[k], s1[k]);
                }
            }
        }
        const int64_t obase = (b * COUT + cg * 8) * (int64_t)NPOS;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float r0 = a0[k] + bias[cg * 8 + k], r1 = a1[k] + bias[cg * 8 + k];
            if (RESID) {
                r0 += idt_in[obase + k * NPOS + p0];
                r1 += idt_in[obase + k * NPOS + p1];
            }
            r0 = fmaxf(r0, 0.f);
            r1 = fmaxf(r1, 0.f);
            if (HEAD) {
                outs[(cg * 8 + k) * NPOS + p0] = r0;
                outs[(cg * 8 + k) * NPOS + p1] = r1;
            } else {
                y[obase + k * NPOS + p0] = r0;
                y[obase + k * NPOS + p1] = r1;
            }
            if (SKIP) {
                idt_out[obase + k * NPOS + p0] = s0[k] + bsk[cg * 8 + k];
                idt_out[obase + k * NPOS + p1] = s1[k] + bsk[cg * 8 + k];
            }
        }
    }
    if (HEAD) {   // AdaptiveAvgPool2d(1) + Flatten + Linear(COUT -> 2), fixed summation order
        __syncthreads();
        float* gap = tile;
        for (int c = tid; c < COUT; c += kCnnThreads) {
            float s = 0.f;
            for (int p = 0; p < NPOS; ++p) s += outs[c * NPOS + p];
            gap[c] = s * (1.0f / NPOS);
        }
        __syncthreads();
        const int warp = tid >> 5, lane = tid & 31;
        if (warp < CD_N_CLASSES) {
            float s = 0.f;
            for (int c = lane; c < COUT; c += 32) s = fmaf(fcw[warp * COUT + c], gap[c], s);
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) logits[b * CD_N_CLASSES + warp] = s + fcb[warp];
        }
    }
}

// per-clip fp32 workspace (floats)
template <int NF>
struct PlanF32 {
    using N = NetShape<NF>;
    static constexpr int64_t X0 = 32 * N::H0 * N::W0, H1 = 64 * N::H1 * N::W1, H2 = 128 * N::H2 * N::W2;
    static constexpr int64_t PER_CLIP = X0 + 3 * H1 + 2 * H2;
    static constexpr size_t smA0 = (size_t)(32 * (N::H0 + 2) * (N::W0 + 2)) * 4;
    static constexpr size_t smB0 = (size_t)(64 * (N::H1 + 2) * (N::W1 + 2)) * 4;
    static constexpr size_t smA1 = (size_t)(64 * (N::H1 + 2) * (N::W1 + 2)) * 4;
    static constexpr size_t smB1 = (size_t)(128 * (N::H2 + 2) * (N::W2 + 2) + 128 * N::H2 * N::W2) * 4;
};
constexpr int64_t kCnnChunk = 4096;

size_t cnn_fp32_ws_bytes(int64_t B) {
    const int64_t n = B < kCnnChunk ? B : kCnnChunk;
    return (size_t)(n * PlanF32<kFeatDD>::PER_CLIP) * sizeof(float);       // the larger of the two shapes
}

template <typename K>
static cudaError_t opt_in_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int NF>
static int init_cnn_fp32_nf(cd_handle* h) {
    using N = NetShape<NF>;
    using P = PlanF32<NF>;
    CD_CUDA(h, opt_in_smem(stem_kernel<NF>, StemF32<NF>::SMEM));
    CD_CUDA(h, opt_in_smem(conv3x3_kernel<32, 64, N::H0, N::W0, 2, true, false, false>, P::smA0));
    CD_CUDA(h, opt_in_smem(conv3x3_kernel<64, 64, N::H1, N::W1, 1, false, true, false>, P::smB0));
    CD_CUDA(h, opt_in_smem(conv3x3_kernel<64, 128, N::H1, N::W1, 2, true, false, false>, P::smA1));
    CD_CUDA(h, opt_in_smem(conv3x3_kernel<128, 128, N::H2, N::W2, 1, false, true, true>, P::smB1));
    return CD_OK;
}

// per-device function attributes (called from cd_create with the device current)
int init_cnn_fp32(cd_handle* h) {
    const int rc = init_cnn_fp32_nf<kFeat>(h);
    return rc != CD_OK ? rc : init_cnn_fp32_nf<kFeatDD>(h);
}

template <int NF>
static int launch_cnn_fp32_nf(cd_handle* h, const float* d_feat, int64_t B, float* d_logits, cudaStream_t s) {
    using N = NetShape<NF>;
    using P = PlanF32<NF>;
    auto kA0 = conv3x3_kernel<32, 64, N::H0, N::W0, 2, true, false, false>;
    auto kB0 = conv3x3_kernel<64, 64, N::H1, N::W1, 1, false, true, false>;
    auto kA1 = conv3x3_kernel<64, 128, N::H1, N::W1, 2, true, false, false>;
    auto kB1 = conv3x3_kernel<128, 128, N::H2, N::W2, 1, false, true, true>;
    const CnnWeightsF32& W = h->w32;
    for (int64_t c0 = 0; c0 < B; c0 += kCnnChunk) {
        const int64_t n = (B - c0) < kCnnChunk ? (B - c0) : kCnnChunk;
        float* x0 = h->d_ws;
        float* h1 = x0 + n * P::X0;
        float* i1 = h1 + n * P::H1;
        float* x1 = i1 + n * P::H1;
        float* h2 = x1 + n * P::H1;
        float* i2 = h2 + n * P::H2;
        const int g = (int)n;
        stem_kernel<NF><<<g, kCnnThreads, StemF32<NF>::SMEM, s>>>(d_feat + c0 * N::FEAT_ELEMS, W.stem_w, W.stem_b, x0);
        kA0<<<g, kCnnThreads, P::smA0, s>>>(x0, W.b0c1_w, W.b0c1_b, W.b0sk_w, W.b0sk_b, nullptr, h1, i1, nullptr, nullptr, nullptr);
        kB0<<<g, kCnnThreads, P::smB0, s>>>(h1, W.b0c2_w, W.b0c2_b, nullptr, nullptr, i1, x1, nullptr, nullptr, nullptr, nullptr);
        kA1<<<g, kCnnThreads, P::smA1, s>>>(x1, W.b1c1_w, W.b1c1_b, W.b1sk_w, W.b1sk_b, nullptr, h2, i2, nullptr, nullptr, nullptr);
        kB1<<<g, kCnnThreads, P::smB1, s>>>(h2, W.b1c2_w, W.b1c2_b, nullptr, nullptr, i2, nullptr, nullptr, W.fc_w, W.fc_b,
                                           d_logits + c0 * CD_N_CLASSES);
        h->launches += 5;
        CD_CUDA(h, cudaGetLastError());
    }
    return CD_OK;
}

int launch_cnn_fp32(cd_handle* h, const float* d_feat, int64_t B, float* d_logits, cudaStream_t s, int nf) {
    if (B == 0) return CD_OK;
    return nf == kFeatDD ? launch_cnn_fp32_nf<kFeatDD>(h, d_feat, B, d_logits, s) : launch_cnn_fp32_nf<kFeat>(h, d_feat, B, d_logits, s);
}

}  // namespace cd
