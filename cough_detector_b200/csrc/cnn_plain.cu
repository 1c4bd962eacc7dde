// The reference's two other model types on CUDA cores, fp32: CoughDetector ("standard", src/model.py:43-143) and
// CoughDetectorSmall ("small", :146-207, the DEFAULT model_type of src/train.py:218 and src/inference.py:146).
//
// They are off the north-star path (the hot path is CoughDetectorResidual, which has the tcgen05
// kernels of cnn_tc.cu); they are here so that a checkpoint written by the reference's stock command line loads and
// runs on the GPU instead of raising.  Eval mode: BatchNorm folded into the convolution that feeds it, Dropout /
// Dropout2d identity.  Both networks are fully convolutional up to their global average pool, so any number of feature
// rows works (90, 103, ...).  Three kernels cover both:
//   conv3x3_pool_kernel   dense 3x3/s1/p1 conv + bias + ReLU + MaxPool2d(2)           (ConvBlock; Small's first block)
//   dw_pw_kernel          depthwise 3x3 + bias, pointwise 1x1 + bias (BN folded) + ReLU [+ MaxPool2d(2)]   (Small's blocks 2-4)
//   gap_fc2_kernel        AdaptiveAvgPool2d(1) + Linear + ReLU + Linear, fixed summation order
#include <math.h>
#include <vector>

#include "cd_common.cuh"

namespace cd {

constexpr int kPlainThreads = 128;
constexpr int kCo = 8;                    // output channels per thread

// in [B][CIN][H][W] -> out [B][COUT][H/2][W/2].  w: [CIN * 9][COUT] (tap = ky * 3 + kx), folded with the BatchNorm.
// grid (pooled-pixel tiles, COUT / 8, B); the CTA's 8 output channels' weights sit in shared memory.
__global__ void __launch_bounds__(kPlainThreads)
conv3x3_pool_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ out, int CIN, int COUT, int H, int W) {
    extern __shared__ float sw[];                                    // [CIN * 9][kCo]
    const int cg = blockIdx.y, b = blockIdx.z, HP = H / 2, WP = W / 2;
    for (int i = threadIdx.x; i < CIN * 9 * kCo; i += kPlainThreads) sw[i] = w[(size_t)(i / kCo) * COUT + cg * kCo + (i % kCo)];
    __syncthreads();
    const int pos = blockIdx.x * kPlainThreads + threadIdx.x;
    if (pos >= HP * WP) return;
    const int py = pos / WP, px = pos - py * WP;
    float acc[4][kCo];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < kCo; ++k) acc[q][k] = 0.f;
    const float* ib = in + (size_t)b * CIN * H * W;
    for (int ci = 0; ci < CIN; ++ci) {
        float p[4][4];                                               // input rows 2py-1 .. 2py+2, cols 2px-1 .. 2px+2
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int y = 2 * py - 1 + r, x = 2 * px - 1 + c;
                p[r][c] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(ib + ((size_t)ci * H + y) * W + x) : 0.f;
            }
        const float* wc = sw + ci * 9 * kCo;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float4 wa = *reinterpret_cast<const float4*>(wc + (ky * 3 + kx) * kCo);
                const float4 wb = *reinterpret_cast<const float4*>(wc + (ky * 3 + kx) * kCo + 4);
                const float wv[kCo] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 2; ++dx)
#pragma unroll
                        for (int k = 0; k < kCo; ++k) acc[dy * 2 + dx][k] = fmaf(p[dy + ky][dx + kx], wv[k], acc[dy * 2 + dx][k]);
            }
    }
#pragma unroll
    for (int k = 0; k < kCo; ++k) {
        const int co = cg * kCo + k;
        const float m = fmaxf(fmaxf(acc[0][k], acc[1][k]), fmaxf(acc[2][k], acc[3][k]));     // ReLU and + bias are monotone
        out[(((size_t)b * COUT + co) * HP + py) * WP + px] = fmaxf(m + bias[co], 0.f);
    }
}

// Depthwise 3x3/p1 (+ its bias) then pointwise 1x1 (BatchNorm folded) + ReLU, POOL: + MaxPool2d(2).
// in [B][C][H][W] -> out [B][COUT][H/2][W/2] (POOL) or [B][COUT][H][W].  dw: [C][9], pw: [C][COUT].
template <bool POOL>
__global__ void __launch_bounds__(kPlainThreads)
dw_pw_kernel(const float* __restrict__ in, const float* __restrict__ dw, const float* __restrict__ dwb,
             const float* __restrict__ pw, const float* __restrict__ pwb, float* __restrict__ out, int C, int COUT, int H, int W) {
    extern __shared__ float sm[];
    float* s_dw = sm;                      // [C][9]
    float* s_db = s_dw + C * 9;            // [C]
    float* s_pw = s_db + ((C + 3) & ~3);   // [C][kCo]
    const int cg = blockIdx.y, b = blockIdx.z;
    const int HO = POOL ? H / 2 : H, WO = POOL ? W / 2 : W;
    for (int i = threadIdx.x; i < C * 9; i += kPlainThreads) s_dw[i] = dw[i];
    for (int i = threadIdx.x; i < C; i += kPlainThreads) s_db[i] = dwb[i];
    for (int i = threadIdx.x; i < C * kCo; i += kPlainThreads) s_pw[i] = pw[(size_t)(i / kCo) * COUT + cg * kCo + (i % kCo)];
    __syncthreads();
    const int pos = blockIdx.x * kPlainThreads + threadIdx.x;
    if (pos >= HO * WO) return;
    const int oy = pos / WO, ox = pos - oy * WO;
    constexpr int NP = POOL ? 4 : 1, PR = POOL ? 4 : 3;
    const int y0 = (POOL ? 2 * oy : oy) - 1, x0 = (POOL ? 2 * ox : ox) - 1;
    float acc[NP][kCo];
#pragma unroll
    for (int q = 0; q < NP; ++q)
#pragma unroll
        for (int k = 0; k < kCo; ++k) acc[q][k] = 0.f;
    const float* ib = in + (size_t)b * C * H * W;
    for (int c = 0; c < C; ++c) {
        float p[PR][PR];
#pragma unroll
        for (int r = 0; r < PR; ++r)
#pragma unroll
            for (int q = 0; q < PR; ++q) {
                const int y = y0 + r, x = x0 + q;
                p[r][q] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(ib + ((size_t)c * H + y) * W + x) : 0.f;
            }
        float d[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int dy = q >> 1, dx = q & 1;
            float s = s_db[c];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) s = fmaf(p[dy + ky][dx + kx], s_dw[c * 9 + ky * 3 + kx], s);
            d[q] = s;
        }
        const float4 wa = *reinterpret_cast<const float4*>(s_pw + c * kCo), wb = *reinterpret_cast<const float4*>(s_pw + c * kCo + 4);
        const float wv[kCo] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int q = 0; q < NP; ++q)
#pragma unroll
            for (int k = 0; k < kCo; ++k) acc[q][k] = fmaf(d[q], wv[k], acc[q][k]);
    }
#pragma unroll
    for (int k = 0; k < kCo; ++k) {
        const int co = cg * kCo + k;
        float m = acc[0][k];
        if (POOL) m = fmaxf(fmaxf(acc[0][k], acc[1][k]), fmaxf(acc[2][k], acc[3][k]));
        out[(((size_t)b * COUT + co) * HO + oy) * WO + ox] = fmaxf(m + pwb[co], 0.f);
    }
}

// x [B][C][P] -> mean over P -> Linear(C -> HID) + ReLU -> Linear(HID -> 2).  One CTA per clip, fixed summation order.
__global__ void __launch_bounds__(256)
gap_fc2_kernel(const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ b1,
               const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ logits, int C, int P, int HID) {
    __shared__ float s_x[256], s_h[256];
    const int b = blockIdx.x, t = threadIdx.x;
    if (t < C) {
        const float* p = x + ((size_t)b * C + t) * P;
        float s = 0.f;
        for (int i = 0; i < P; ++i) s += p[i];
        s_x[t] = s / (float)P;
    }
    __syncthreads();
    if (t < HID) {
        float s = b1[t];
        for (int c = 0; c < C; ++c) s = fmaf(w1[(size_t)t * C + c], s_x[c], s);
        s_h[t] = fmaxf(s, 0.f);
    }
    __syncthreads();
    if (t < CD_N_CLASSES) {
        float s = b2[t];
        for (int j = 0; j < HID; ++j) s = fmaf(w2[(size_t)t * HID + j], s_h[j], s);
        logits[(size_t)b * CD_N_CLASSES + t] = s;
    }
}

// ----------------------------------------------------------------------------------------------
// host side: a network is a list of layers over one device blob
// ----------------------------------------------------------------------------------------------
struct PlainLayer {
    int kind;                  // 0 conv3x3+pool, 1 dw+pw+pool, 2 dw+pw, 3 gap+fc2
    int cin, cout, hid;
    size_t w, b, w2, b2;       // offsets (floats) into the blob: conv w/b | dw w/b + pw w2/b2 | fc1 w/b + fc2 w2/b2
};
struct PlainNet {
    std::vector<PlainLayer> layers;
    float* d_blob = nullptr;
};

static double bn_scale(const cd_bn& bn, int c) { return (double)bn.weight[c] / sqrt((double)bn.running_var[c] + 1e-5); }

void free_plain(cd_handle* h) {
    for (int i = 0; i < 2; ++i)
        if (h->plain[i]) {
            cudaFree(h->plain[i]->d_blob);
            delete h->plain[i];
            h->plain[i] = nullptr;
        }
}

static size_t push(std::vector<float>& blob, const std::vector<float>& v) {
    const size_t off = blob.size();
    blob.insert(blob.end(), v.begin(), v.end());
    while (blob.size() % 4) blob.push_back(0.f);
    return off;
}

// conv [cout][cin][3][3] + BN -> w [cin * 9 + tap][cout], b [cout]
static void fold_conv3(const cd_conv& cv, const cd_bn& bn, int cin, int cout, std::vector<float>& w, std::vector<float>& b) {
    w.assign((size_t)cin * 9 * cout, 0.f);
    b.assign(cout, 0.f);
    for (int co = 0; co < cout; ++co) {
        const double s = bn_scale(bn, co);
        b[co] = (float)(((double)cv.bias[co] - (double)bn.running_mean[co]) * s + (double)bn.bias[co]);
        for (int ci = 0; ci < cin; ++ci)
            for (int t = 0; t < 9; ++t) w[((size_t)ci * 9 + t) * cout + co] = (float)((double)cv.weight[((size_t)co * cin + ci) * 9 + t] * s);
    }
}

static int upload_plain(cd_handle* h, int slot, PlainNet* net, const std::vector<float>& blob) {
    cudaError_t e = cudaMalloc(&net->d_blob, blob.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(net->d_blob, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(net->d_blob);
        delete net;
        return set_err(h, CD_E_CUDA, "plain network upload", e);
    }
    CD_CUDA(h, cudaDeviceSynchronize());            // earlier launches may still read the old blob
    if (h->plain[slot]) {
        cudaFree(h->plain[slot]->d_blob);
        delete h->plain[slot];
    }
    h->plain[slot] = net;
    return CD_OK;
}

static bool ok(const cd_conv& c) { return c.weight && c.bias; }
static bool ok(const cd_bn& b) { return b.weight && b.bias && b.running_mean && b.running_var; }

int load_small(cd_handle* h, const cd_small_weights* w) {
    bool good = ok(w->conv1) && ok(w->bn1) && w->fc1_weight && w->fc1_bias && w->fc2_weight && w->fc2_bias;
    for (int i = 0; i < 3; ++i) good = good && ok(w->block[i].dw) && ok(w->block[i].pw) && ok(w->block[i].bn);
    if (!good) return set_err(h, CD_E_ARG, "cd_small_load: a tensor pointer is NULL");
    PlainNet* net = new PlainNet();
    std::vector<float> blob, a, b;
    fold_conv3(w->conv1, w->bn1, 1, 16, a, b);
    PlainLayer L0{0, 1, 16, 0, push(blob, a), 0, 0, 0};
    L0.b = push(blob, b);
    net->layers.push_back(L0);
    const int ch[4] = {16, 32, 64, 128};
    for (int i = 0; i < 3; ++i) {
        const int c = ch[i], co = ch[i + 1];
        std::vector<float> dw(w->block[i].dw.weight, w->block[i].dw.weight + (size_t)c * 9), db(w->block[i].dw.bias, w->block[i].dw.bias + c);
        std::vector<float> pw((size_t)c * co), pb(co);
        for (int o = 0; o < co; ++o) {
            const double s = bn_scale(w->block[i].bn, o);
            pb[o] = (float)(((double)w->block[i].pw.bias[o] - (double)w->block[i].bn.running_mean[o]) * s + (double)w->block[i].bn.bias[o]);
            for (int k = 0; k < c; ++k) pw[(size_t)k * co + o] = (float)((double)w->block[i].pw.weight[(size_t)o * c + k] * s);
        }
        PlainLayer L{i < 2 ? 1 : 2, c, co, 0, 0, 0, 0, 0};
        L.w = push(blob, dw); L.b = push(blob, db); L.w2 = push(blob, pw); L.b2 = push(blob, pb);
        net->layers.push_back(L);
    }
    PlainLayer F{3, 128, CD_N_CLASSES, 64, 0, 0, 0, 0};
    F.w = push(blob, std::vector<float>(w->fc1_weight, w->fc1_weight + 64 * 128));
    F.b = push(blob, std::vector<float>(w->fc1_bias, w->fc1_bias + 64));
    F.w2 = push(blob, std::vector<float>(w->fc2_weight, w->fc2_weight + CD_N_CLASSES * 64));
    F.b2 = push(blob, std::vector<float>(w->fc2_bias, w->fc2_bias + CD_N_CLASSES));
    net->layers.push_back(F);
    return upload_plain(h, 0, net, blob);
}

int load_standard(cd_handle* h, const cd_standard_weights* w) {
    bool good = w->fc1_weight && w->fc1_bias && w->fc2_weight && w->fc2_bias;
    for (int i = 0; i < 4; ++i) good = good && ok(w->conv[i]) && ok(w->bn[i]);
    if (!good) return set_err(h, CD_E_ARG, "cd_standard_load: a tensor pointer is NULL");
    PlainNet* net = new PlainNet();
    std::vector<float> blob, a, b;
    const int ch[5] = {1, 32, 64, 128, 256};
    for (int i = 0; i < 4; ++i) {
        fold_conv3(w->conv[i], w->bn[i], ch[i], ch[i + 1], a, b);
        PlainLayer L{0, ch[i], ch[i + 1], 0, push(blob, a), 0, 0, 0};
        L.b = push(blob, b);
        net->layers.push_back(L);
    }
    PlainLayer F{3, 256, CD_N_CLASSES, 128, 0, 0, 0, 0};
    F.w = push(blob, std::vector<float>(w->fc1_weight, w->fc1_weight + 128 * 256));
    F.b = push(blob, std::vector<float>(w->fc1_bias, w->fc1_bias + 128));
    F.w2 = push(blob, std::vector<float>(w->fc2_weight, w->fc2_weight + CD_N_CLASSES * 128));
    F.b2 = push(blob, std::vector<float>(w->fc2_bias, w->fc2_bias + CD_N_CLASSES));
    net->layers.push_back(F);
    return upload_plain(h, 1, net, blob);
}

// largest activation (floats per clip) any layer writes, for n_rows x 101 inputs
static size_t plain_act_elems(const PlainNet& net, int H, int W) {
    size_t m = 0;
    for (const PlainLayer& L : net.layers) {
        if (L.kind == 0 || L.kind == 1) { H /= 2; W /= 2; }
        if (L.kind != 3) m = std::max(m, (size_t)L.cout * H * W);
    }
    return m;
}

int init_cnn_plain(cd_handle* h) {
    CD_CUDA(h, cudaFuncSetAttribute(conv3x3_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 9 * kCo * 4));
    return CD_OK;
}

int launch_plain(cd_handle* h, int slot, const float* d_feat, int n_rows, int64_t B, float* d_logits, cudaStream_t s) {
    const PlainNet* net = h->plain[slot];
    if (!net) return set_err(h, CD_E_STATE, slot ? "cd_standard_forward: call cd_standard_load first" : "cd_small_forward: call cd_small_load first");
    if (B == 0) return CD_OK;
    const size_t act = plain_act_elems(*net, n_rows, kFrames);
    const int64_t chunk = 1024;                                        // clips per pass through the layers (bounds the workspace)
    int rc = ensure_ws(h, 2 * act * (size_t)chunk * sizeof(float));
    if (rc != CD_OK) return rc;
    float* buf[2] = {h->d_ws, h->d_ws + act * chunk};
    for (int64_t b0 = 0; b0 < B; b0 += chunk) {
        const int n = (int)(B - b0 < chunk ? B - b0 : chunk);
        const float* src = d_feat + b0 * (int64_t)n_rows * kFrames;
        int H = n_rows, W = kFrames, cur = 0;
        for (const PlainLayer& L : net->layers) {
            const float* blob = net->d_blob;
            float* dst = buf[cur];
            if (L.kind == 0) {
                dim3 grid((unsigned)(((H / 2) * (W / 2) + kPlainThreads - 1) / kPlainThreads), (unsigned)(L.cout / kCo), (unsigned)n);
                conv3x3_pool_kernel<<<grid, kPlainThreads, (size_t)L.cin * 9 * kCo * sizeof(float), s>>>(src, blob + L.w, blob + L.b, dst, L.cin, L.cout, H, W);
                H /= 2; W /= 2;
            } else if (L.kind == 1 || L.kind == 2) {
                const bool pool = L.kind == 1;
                const int HO = pool ? H / 2 : H, WO = pool ? W / 2 : W;
                dim3 grid((unsigned)((HO * WO + kPlainThreads - 1) / kPlainThreads), (unsigned)(L.cout / kCo), (unsigned)n);
                const size_t smem = ((size_t)L.cin * 9 + ((L.cin + 3) & ~3) + (size_t)L.cin * kCo) * sizeof(float);
                if (pool) dw_pw_kernel<true><<<grid, kPlainThreads, smem, s>>>(src, blob + L.w, blob + L.b, blob + L.w2, blob + L.b2, dst, L.cin, L.cout, H, W);
                else dw_pw_kernel<false><<<grid, kPlainThreads, smem, s>>>(src, blob + L.w, blob + L.b, blob + L.w2, blob + L.b2, dst, L.cin, L.cout, H, W);
                H = HO; W = WO;
            } else {
                gap_fc2_kernel<<<n, 256, 0, s>>>(src, blob + L.w, blob + L.b, blob + L.w2, blob + L.b2, d_logits + b0 * CD_N_CLASSES, L.cin, H * W, L.hid);
            }
            h->launches += 1;
            src = dst;
            cur ^= 1;
        }
        CD_CUDA(h, cudaGetLastError());
    }
    return CD_OK;
}

}  // namespace cd
