// C ABI of libcough_b200 (include/cough_b200.h): handle lifetime, BatchNorm folding + weight
// upload, argument checking, and the host-buffer pipeline.  Kernels live in feat_kernel.cu,
// cnn_fp32.cu and cnn_tc.cu.
#include <math.h>
#include <mutex>
#include <vector>

#include "cd_common.cuh"
#include <stdlib.h>

namespace cd {

static thread_local std::string g_create_err;

int set_err(cd_handle* h, int code, const char* what, cudaError_t e) {
    std::string msg = what ? what : "";
    if (e != cudaSuccess) {
        msg += ": ";
        msg += cudaGetErrorString(e);
    }
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}

int ensure_ws(cd_handle* h, size_t bytes) {
    if (bytes <= h->ws_bytes) return CD_OK;
    if (h->d_ws) {
        CD_CUDA(h, cudaDeviceSynchronize());     // earlier launches may still use the old buffer
        cudaFree(h->d_ws);
        h->d_ws = nullptr;
        h->ws_bytes = 0;
    }
    cudaError_t e = cudaMalloc(&h->d_ws, bytes);
    if (e != cudaSuccess) return set_err(h, CD_E_NOMEM, "workspace allocation", e);
    h->ws_bytes = bytes;
    return CD_OK;
}

// Successive users of the handle's workspaces on different streams are ordered with one event (cd_common.cuh).
static int ws_acquire(cd_handle* h, cudaStream_t s) {
    if (h->ws_busy && h->ws_stream != s) CD_CUDA(h, cudaStreamWaitEvent(s, h->ws_event, 0));
    return CD_OK;
}
static int ws_release(cd_handle* h, cudaStream_t s) {
    CD_CUDA(h, cudaEventRecord(h->ws_event, s));
    h->ws_stream = s;
    h->ws_busy = true;
    return CD_OK;
}

int build_feat_tables(cd_handle* h);
int init_cnn_fp32(cd_handle* h);
int init_cnn_tc(cd_handle* h);
int upload_cnn_tc(cd_handle* h, const std::vector<std::vector<float>>& wf, const std::vector<std::vector<float>>& bf);
int launch_cnn_tc(cd_handle* h, const float* d_feat, int64_t B, float* d_logits, cudaStream_t s, int nf = kFeat);
size_t cnn_tc_ws_bytes(int64_t B);
__nv_bfloat16* cnn_tc_map_ptr(cd_handle* h);
int ensure_tc_ws(cd_handle* h, int nf = kFeat);
int launch_features_i16(cd_handle* h, const int16_t* d_audio, int64_t stride, int mode, float coef, int64_t B,
                        float* d_feat, __nv_bfloat16* d_map, cudaStream_t s);
void free_resample_tables(cd_handle* h);
void free_plain(cd_handle* h);
int init_cnn_plain(cd_handle* h);
int load_small(cd_handle* h, const cd_small_weights* w);
int load_standard(cd_handle* h, const cd_standard_weights* w);
int launch_plain(cd_handle* h, int slot, const float* d_feat, int n_rows, int64_t B, float* d_logits, cudaStream_t s);
int launch_condition_var(cd_handle* h, const void* d_in, bool i16, const int64_t* d_in_off, const int32_t* d_chans,
                         const int32_t* d_n_in, const int64_t* d_out_off, int64_t B, int64_t max_n_in, int orig_sr, int new_sr,
                         float* d_out, cudaStream_t s);
int64_t resample_length(int64_t n_in, int orig_sr, int new_sr);
int launch_resample(cd_handle* h, const float* d_in, int64_t rows, int64_t n_in, int orig_sr, int new_sr, float* d_out,
                    cudaStream_t s);
void get_host_tables(float* hann, float* fb, float* dct);
int umma_selftest(cd_handle* h, int N, int K, int shift, int mode, const float* hA, const float* hB, float* hD);

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

__global__ void predict_kernel(const float* __restrict__ logits, int64_t B, float thr, float* __restrict__ probs,
                               int64_t* __restrict__ preds, uint8_t* __restrict__ dec) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float l0 = logits[2 * i], l1 = logits[2 * i + 1];
    const float m = fmaxf(l0, l1);
    const float e0 = expf(l0 - m), e1 = expf(l1 - m);
    const float inv = 1.0f / (e0 + e1);
    const float p0 = e0 * inv, p1 = e1 * inv;
    if (probs) {
        probs[2 * i] = p0;
        probs[2 * i + 1] = p1;
    }
    if (preds) preds[i] = p1 > p0 ? 1 : 0;          // argmax, first index wins ties (torch.argmax)
    if (dec) dec[i] = p1 >= thr ? 1 : 0;
}

}  // namespace cd

using namespace cd;

extern "C" {

const char* cd_version(void) { return "cough_b200 0.1 sm_100a"; }

const char* cd_last_error(const cd_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int64_t cd_launch_count(const cd_handle* h) { return h ? h->launches : 0; }

int cd_selftest_umma(cd_handle* h, int N, int K, int shift, int mode, const float* h_A, const float* h_B, float* h_D) {
    if (!h) return CD_E_ARG;
    if (!h_A || !h_B || !h_D) return set_err(h, CD_E_ARG, "cd_selftest_umma: NULL pointer");
    DeviceGuard g(h->device);
    return umma_selftest(h, N, K, shift, mode, h_A, h_B, h_D);
}

int cd_get_tables(float* hann, float* fb, float* dct) {
    get_host_tables(hann, fb, dct);
    return CD_OK;
}

int64_t cd_stream_windows(int64_t n_samples, int64_t hop) {
    if (hop <= 0 || n_samples < kSeg) return 0;
    return (n_samples - kSeg) / hop + 1;
}

int cd_create(int device, cd_handle** out) {
    if (!out) return set_err(nullptr, CD_E_ARG, "cd_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return set_err(nullptr, CD_E_NODEVICE, "no CUDA device", e);
    if (device < 0 || device >= n) return set_err(nullptr, CD_E_ARG, "cd_create: device index out of range");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return set_err(nullptr, CD_E_CUDA, "cudaGetDeviceProperties", e);
    if (prop.major != 10) {
        char buf[160];
        snprintf(buf, sizeof buf, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                 prop.major, prop.minor);
        return set_err(nullptr, CD_E_NODEVICE, buf);
    }
    DeviceGuard g(device);
    cd_handle* h = new cd_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    int rc = build_feat_tables(h);
    if (rc == CD_OK && cudaEventCreateWithFlags(&h->ws_event, cudaEventDisableTiming) != cudaSuccess)
        rc = set_err(h, CD_E_CUDA, "cudaEventCreate");
    if (rc == CD_OK) rc = init_cnn_fp32(h);
    if (rc == CD_OK) rc = init_cnn_tc(h);
    if (rc == CD_OK) rc = init_cnn_plain(h);
    if (rc != CD_OK) {
        g_create_err = h->err;
        cd_destroy(h);
        return rc;
    }
    *out = h;
    return CD_OK;
}

int cd_destroy(cd_handle* h) {
    if (!h) return CD_OK;
    DeviceGuard g(h->device);
    cudaDeviceSynchronize();
    cudaFree(h->d_tables);
    cudaFree(h->d_wblob);
    cudaFree(h->d_wblob_bf16);
    cudaFree(h->d_ws_tc_fam[0]);
    cudaFree(h->d_ws_tc_fam[1]);
    cudaFree(h->d_ws_stream);
    cudaFree(h->d_ws);
    for (int i = 0; i < 3; ++i) {
        if (h->pipe_streams[i]) cudaStreamDestroy(h->pipe_streams[i]);
        if (h->pipe_events[i]) cudaEventDestroy(h->pipe_events[i]);
        cudaFree(h->dev_in[i]);
    }
    cudaFree(h->dev_logits);
    free_resample_tables(h);
    free_plain(h);
    if (h->ws_event) cudaEventDestroy(h->ws_event);
    delete h;
    return CD_OK;
}

// ---- features --------------------------------------------------------------------------------

// CD_FEAT_* flags of a call -> FeatMode bits of the kernels; nothing is taken from the handle.
static bool feat_flags_ok(cd_handle* h, int flags, float coef, const char* who) {
    if (flags & ~(CD_FEAT_NORMALIZE | CD_FEAT_PRE_EMPHASIS | CD_FEAT_PCEN | CD_FEAT_DELTA_DELTA)) {
        set_err(h, CD_E_ARG, (std::string(who) + ": flags are CD_FEAT_NORMALIZE | CD_FEAT_PRE_EMPHASIS | CD_FEAT_PCEN | CD_FEAT_DELTA_DELTA").c_str());
        return false;
    }
    if ((flags & CD_FEAT_PRE_EMPHASIS) && !(coef >= 0.f && coef <= 1.f)) {
        set_err(h, CD_E_ARG, (std::string(who) + ": pre_emphasis_coef must be in [0, 1]").c_str());
        return false;
    }
    return true;
}
static int feat_mode_bits(int flags) {
    return ((flags & CD_FEAT_NORMALIZE) ? 0 : kModeNoNormalize) | ((flags & CD_FEAT_PRE_EMPHASIS) ? kModePreEmph : 0) |
           ((flags & CD_FEAT_PCEN) ? kModePcen : 0) | ((flags & CD_FEAT_DELTA_DELTA) ? kModeDeltaDelta : 0);
}

int cd_feat_forward(cd_handle* h, const float* d_audio, int64_t stride, int64_t B, float* d_feat, int flags,
                    float pre_emphasis_coef, void* stream) {
    if (!h) return CD_E_ARG;
    if (B < 0 || stride < kSeg || (B > 0 && (!d_audio || !d_feat)))
        return set_err(h, CD_E_ARG, "cd_feat_forward: bad argument (need stride >= 16000, non-NULL buffers)");
    if (!feat_flags_ok(h, flags, pre_emphasis_coef, "cd_feat_forward")) return CD_E_ARG;
    DeviceGuard g(h->device);
    return launch_features(h, d_audio, nullptr, nullptr, stride, kModeFixed | feat_mode_bits(flags), pre_emphasis_coef, B,
                           d_feat, nullptr, (cudaStream_t)stream);
}

int cd_feat_forward_i16(cd_handle* h, const int16_t* d_audio, int64_t stride, int64_t B, float* d_feat, int flags,
                        float pre_emphasis_coef, void* stream) {
    if (!h) return CD_E_ARG;
    if (B < 0 || stride < kSeg || (stride & 7) || (B > 0 && (!d_audio || !d_feat)))
        return set_err(h, CD_E_ARG, "cd_feat_forward_i16: bad argument (need stride >= 16000 and a multiple of 8)");
    if (!feat_flags_ok(h, flags, pre_emphasis_coef, "cd_feat_forward_i16")) return CD_E_ARG;
    DeviceGuard g(h->device);
    return launch_features_i16(h, d_audio, stride, kModeFixed | feat_mode_bits(flags), pre_emphasis_coef, B, d_feat, nullptr,
                               (cudaStream_t)stream);
}

int cd_feat_forward_var(cd_handle* h, const float* d_audio, const int64_t* d_offsets, const int32_t* d_lengths,
                        int64_t B, float* d_feat, int flags, float pre_emphasis_coef, void* stream) {
    if (!h) return CD_E_ARG;
    if (B < 0 || (B > 0 && (!d_audio || !d_offsets || !d_lengths || !d_feat)))
        return set_err(h, CD_E_ARG, "cd_feat_forward_var: bad argument");
    if (!feat_flags_ok(h, flags, pre_emphasis_coef, "cd_feat_forward_var")) return CD_E_ARG;
    DeviceGuard g(h->device);
    return launch_features(h, d_audio, d_offsets, d_lengths, 0, kModeVar | feat_mode_bits(flags | CD_FEAT_NORMALIZE),
                           pre_emphasis_coef, B, d_feat, nullptr, (cudaStream_t)stream);
}

int cd_feat_forward_stream(cd_handle* h, const float* d_audio, int64_t n_samples, int64_t hop, int64_t n_windows,
                           float* d_feat, int flags, float pre_emphasis_coef, void* stream) {
    if (!h) return CD_E_ARG;
    if (hop <= 0 || n_windows != cd_stream_windows(n_samples, hop) || (n_windows > 0 && (!d_audio || !d_feat)))
        return set_err(h, CD_E_ARG, "cd_feat_forward_stream: hop must be > 0 and n_windows == cd_stream_windows()");
    if (!feat_flags_ok(h, flags, pre_emphasis_coef, "cd_feat_forward_stream")) return CD_E_ARG;
    DeviceGuard g(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ws_acquire(h, s);                               // the shared-STFT path keeps its frames in d_ws_stream
    if (rc == CD_OK)
        rc = launch_features(h, d_audio, nullptr, nullptr, hop, kModeStream | feat_mode_bits(flags | CD_FEAT_NORMALIZE),
                             pre_emphasis_coef, n_windows, d_feat, nullptr, s);
    return rc == CD_OK ? ws_release(h, s) : rc;
}

// ---- resampling ------------------------------------------------------------------------------

int64_t cd_resample_length(int64_t n_in, int orig_sr, int new_sr) {
    if (n_in < 0 || orig_sr <= 0 || new_sr <= 0) return 0;
    return resample_length(n_in, orig_sr, new_sr);
}

int cd_resample(cd_handle* h, const float* d_in, int64_t rows, int64_t n_in, int orig_sr, int new_sr, float* d_out,
                void* stream) {
    if (!h) return CD_E_ARG;
    if (rows < 0 || n_in < 0 || orig_sr <= 0 || new_sr <= 0 || (rows * n_in > 0 && (!d_in || !d_out)))
        return set_err(h, CD_E_ARG, "cd_resample: bad argument");
    DeviceGuard g(h->device);
    return launch_resample(h, d_in, rows, n_in, orig_sr, new_sr, d_out, (cudaStream_t)stream);
}

int cd_condition_var(cd_handle* h, const void* d_in, int sample_type, const int64_t* d_in_offsets, const int32_t* d_channels,
                     const int32_t* d_n_in, const int64_t* d_out_offsets, int64_t B, int64_t max_n_in, int orig_sr, int new_sr,
                     float* d_out, void* stream) {
    if (!h) return CD_E_ARG;
    if (B < 0 || max_n_in < 0 || orig_sr <= 0 || new_sr <= 0 || (sample_type != 0 && sample_type != 1) ||
        (B > 0 && (!d_in || !d_in_offsets || !d_channels || !d_n_in || !d_out_offsets || !d_out)))
        return set_err(h, CD_E_ARG, "cd_condition_var: bad argument");
    DeviceGuard g(h->device);
    return launch_condition_var(h, d_in, sample_type == 1, d_in_offsets, d_channels, d_n_in, d_out_offsets, B, max_n_in, orig_sr,
                                new_sr, d_out, (cudaStream_t)stream);
}

int cd_host_alloc(void** out, size_t bytes, int flags) {
    if (!out || (flags != CD_HOST_PINNED && flags != CD_HOST_WRITE_COMBINED)) return CD_E_ARG;
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes, flags == CD_HOST_WRITE_COMBINED ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
    return e == cudaSuccess ? CD_OK : set_err(nullptr, CD_E_NOMEM, "cudaHostAlloc", e);
}

int cd_host_free(void* p) {
    if (!p) return CD_OK;
    return cudaFreeHost(p) == cudaSuccess ? CD_OK : CD_E_CUDA;
}

// ---- model -----------------------------------------------------------------------------------

// w' = w * g / sqrt(v + eps), b' = (b - mu) * g / sqrt(v + eps) + beta   (eval-mode BatchNorm2d, eps 1e-5)
static bool fold(const cd_conv_bn& c, int cout, int cin, int k, std::vector<float>& wf, std::vector<float>& bf) {
    if (!c.weight || !c.bias || !c.bn_weight || !c.bn_bias || !c.bn_running_mean || !c.bn_running_var) return false;
    const int kk = k * k;
    wf.assign((size_t)kk * cin * cout, 0.f);        // [tap][cin][cout]
    bf.assign(cout, 0.f);
    for (int co = 0; co < cout; ++co) {
        const double s = (double)c.bn_weight[co] / sqrt((double)c.bn_running_var[co] + 1e-5);
        bf[co] = (float)(((double)c.bias[co] - (double)c.bn_running_mean[co]) * s + (double)c.bn_bias[co]);
        for (int ci = 0; ci < cin; ++ci)
            for (int t = 0; t < kk; ++t)
                wf[((size_t)t * cin + ci) * cout + co] = (float)((double)c.weight[((size_t)co * cin + ci) * kk + t] * s);
    }
    return true;
}

int cd_cnn_load(cd_handle* h, const cd_residual_weights* w) {
    if (!h) return CD_E_ARG;
    if (!w || !w->fc_weight || !w->fc_bias) return set_err(h, CD_E_ARG, "cd_cnn_load: NULL weights");
    DeviceGuard g(h->device);
    std::vector<std::vector<float>> wf(8), bf(8);
    bool ok = fold(w->stem, 32, 1, 7, wf[0], bf[0]) && fold(w->b0_conv1, 64, 32, 3, wf[1], bf[1]) &&
              fold(w->b0_skip, 64, 32, 1, wf[2], bf[2]) && fold(w->b0_conv2, 64, 64, 3, wf[3], bf[3]) &&
              fold(w->b1_conv1, 128, 64, 3, wf[4], bf[4]) && fold(w->b1_skip, 128, 64, 1, wf[5], bf[5]) &&
              fold(w->b1_conv2, 128, 128, 3, wf[6], bf[6]);
    if (!ok) return set_err(h, CD_E_ARG, "cd_cnn_load: a conv/bn tensor pointer is NULL");
    wf[7].assign(w->fc_weight, w->fc_weight + 2 * 128);
    bf[7].assign(w->fc_bias, w->fc_bias + 2);
    size_t total = 0;
    for (int i = 0; i < 8; ++i) total += ((wf[i].size() + 3) & ~size_t(3)) + ((bf[i].size() + 3) & ~size_t(3));
    std::vector<float> blob(total, 0.f);
    CD_CUDA(h, cudaDeviceSynchronize());
    if (!h->d_wblob) CD_CUDA(h, cudaMalloc(&h->d_wblob, total * sizeof(float)));
    float** slots[16] = {&h->w32.stem_w, &h->w32.stem_b, &h->w32.b0c1_w, &h->w32.b0c1_b, &h->w32.b0sk_w, &h->w32.b0sk_b,
                         &h->w32.b0c2_w, &h->w32.b0c2_b, &h->w32.b1c1_w, &h->w32.b1c1_b, &h->w32.b1sk_w, &h->w32.b1sk_b,
                         &h->w32.b1c2_w, &h->w32.b1c2_b, &h->w32.fc_w, &h->w32.fc_b};
    size_t off = 0;
    for (int i = 0; i < 8; ++i) {
        memcpy(blob.data() + off, wf[i].data(), wf[i].size() * 4);
        *slots[2 * i] = h->d_wblob + off;
        off += (wf[i].size() + 3) & ~size_t(3);
        memcpy(blob.data() + off, bf[i].data(), bf[i].size() * 4);
        *slots[2 * i + 1] = h->d_wblob + off;
        off += (bf[i].size() + 3) & ~size_t(3);
    }
    CD_CUDA(h, cudaMemcpy(h->d_wblob, blob.data(), total * sizeof(float), cudaMemcpyHostToDevice));
    int rc = upload_cnn_tc(h, wf, bf);
    if (rc != CD_OK) return rc;
    h->weights_loaded = true;
    return CD_OK;
}

static int cnn_forward_impl(cd_handle* h, const float* d_feat, int n_rows, int64_t B, float* d_logits, int precision,
                            cudaStream_t s) {
    if (!h->weights_loaded) return set_err(h, CD_E_STATE, "cd_cnn_forward: call cd_cnn_load first");
    if (n_rows != kFeat && n_rows != kFeatDD)
        return set_err(h, CD_E_ARG, "cd_cnn_forward: the network is built for 90 or 103 (use_delta_delta) feature rows");
    if (precision != CD_PRECISION_FP32 && precision != CD_PRECISION_BF16)
        return set_err(h, CD_E_ARG, "precision must be CD_PRECISION_FP32 or CD_PRECISION_BF16");
    if (B == 0) return CD_OK;
    const bool tcp = precision == CD_PRECISION_BF16;
    int rc = ensure_ws(h, tcp ? cnn_tc_ws_bytes(B) : cnn_fp32_ws_bytes(B));
    if (rc == CD_OK) rc = ws_acquire(h, s);
    if (rc == CD_OK) rc = tcp ? launch_cnn_tc(h, d_feat, B, d_logits, s, n_rows) : launch_cnn_fp32(h, d_feat, B, d_logits, s, n_rows);
    return rc == CD_OK ? ws_release(h, s) : rc;
}

int cd_cnn_forward(cd_handle* h, const float* d_feat, int64_t B, float* d_logits, int precision, void* stream) {
    if (!h) return CD_E_ARG;
    if (B < 0 || (B > 0 && (!d_feat || !d_logits))) return set_err(h, CD_E_ARG, "cd_cnn_forward: bad argument");
    DeviceGuard g(h->device);
    return cnn_forward_impl(h, d_feat, kFeat, B, d_logits, precision, (cudaStream_t)stream);
}

int cd_cnn_forward_rows(cd_handle* h, const float* d_feat, int n_rows, int64_t B, float* d_logits, int precision,
                        void* stream) {
    if (!h) return CD_E_ARG;
    if (B < 0 || (B > 0 && (!d_feat || !d_logits))) return set_err(h, CD_E_ARG, "cd_cnn_forward_rows: bad argument");
    DeviceGuard g(h->device);
    return cnn_forward_impl(h, d_feat, n_rows, B, d_logits, precision, (cudaStream_t)stream);
}

static int plain_forward(cd_handle* h, int slot, const float* d_feat, int n_rows, int64_t B, float* d_logits, void* stream) {
    if (!h) return CD_E_ARG;
    if (B < 0 || n_rows < 8 || n_rows > 4096 || (B > 0 && (!d_feat || !d_logits)))
        return set_err(h, CD_E_ARG, "cd_small_forward / cd_standard_forward: bad argument (need n_rows >= 8)");
    DeviceGuard g(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ws_acquire(h, s);
    if (rc == CD_OK) rc = launch_plain(h, slot, d_feat, n_rows, B, d_logits, s);
    return rc == CD_OK ? ws_release(h, s) : rc;
}

int cd_small_load(cd_handle* h, const cd_small_weights* w) {
    if (!h) return CD_E_ARG;
    if (!w) return set_err(h, CD_E_ARG, "cd_small_load: NULL weights");
    DeviceGuard g(h->device);
    return load_small(h, w);
}
int cd_small_forward(cd_handle* h, const float* d_feat, int n_rows, int64_t B, float* d_logits, void* stream) {
    return plain_forward(h, 0, d_feat, n_rows, B, d_logits, stream);
}
int cd_standard_load(cd_handle* h, const cd_standard_weights* w) {
    if (!h) return CD_E_ARG;
    if (!w) return set_err(h, CD_E_ARG, "cd_standard_load: NULL weights");
    DeviceGuard g(h->device);
    return load_standard(h, w);
}
int cd_standard_forward(cd_handle* h, const float* d_feat, int n_rows, int64_t B, float* d_logits, void* stream) {
    return plain_forward(h, 1, d_feat, n_rows, B, d_logits, stream);
}

int cd_predict(cd_handle* h, const float* d_logits, int64_t B, float threshold, float* d_probs, int64_t* d_preds,
               uint8_t* d_decisions, void* stream) {
    if (!h) return CD_E_ARG;
    if (B < 0 || (B > 0 && !d_logits)) return set_err(h, CD_E_ARG, "cd_predict: bad argument");
    if (B == 0) return CD_OK;
    DeviceGuard g(h->device);
    predict_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_logits, B, threshold, d_probs,
                                                                                 d_preds, d_decisions);
    h->launches += 1;
    CD_CUDA(h, cudaGetLastError());
    return CD_OK;
}

int cd_spec_augment(cd_handle* h, float* d_feat, int n_rows, int64_t B, const int32_t* d_masks, int n_freq_masks,
                    int n_time_masks, void* stream) {
    if (!h) return CD_E_ARG;
    if (B < 0 || n_rows <= 0 || (B > 0 && (!d_feat || (!d_masks && n_freq_masks + n_time_masks > 0))))
        return set_err(h, CD_E_ARG, "cd_spec_augment: bad argument");
    DeviceGuard g(h->device);
    return launch_spec_augment(h, d_feat, n_rows, kFrames, B, d_masks, n_freq_masks, n_time_masks, (cudaStream_t)stream);
}

// ---- end to end ------------------------------------------------------------------------------

// Features for the CNN live in a second, separately grown buffer so the CNN workspace can be
// re-used chunk by chunk.  `mode`: FeatMode bits of this call (entry kind + options).
static int e2e_impl(cd_handle* h, const void* d_audio_v, bool i16, int64_t stride_or_hop, int mode, float coef, int64_t B,
                    float* d_logits, float* d_feat_opt, int precision, cudaStream_t s) {
    const float* d_audio = static_cast<const float*>(d_audio_v);
    const int16_t* d_audio16 = static_cast<const int16_t*>(d_audio_v);
    if (!h->weights_loaded) return set_err(h, CD_E_STATE, "cd_e2e_forward: call cd_cnn_load first");
    if (B == 0) return CD_OK;
    const int nf = (mode & kModeDeltaDelta) ? kFeatDD : kFeat;              // shape family of this call's features
    const int64_t felems = (int64_t)nf * kFrames;
    // process in slabs so a slab's features stay L2-resident between the two stages.  bf16: the
    // feature kernel writes the stem's bf16 map straight into the CNN workspace (fp32 features only
    // if the caller asked for them); fp32: features go through an fp32 buffer.
    int64_t slab = 16384;
    if (const char* e = getenv("CD_E2E_SLAB")) {          // experiment knob (tools/slab_sweep.py): smaller slabs keep the
        const long v = atol(e);                           // activations in L2 but pay the per-launch ramp more often
        if (v >= 148 && v <= 16384) slab = v;
    }
    const bool tcp = precision == CD_PRECISION_BF16;
    const size_t cnn_ws = tcp ? cnn_tc_ws_bytes(slab) : cnn_fp32_ws_bytes(slab);
    const size_t feat_bytes = (tcp || d_feat_opt) ? 0 : (size_t)slab * felems * sizeof(float);
    int rc = ensure_ws(h, cnn_ws + feat_bytes);
    if (rc == CD_OK && tcp) rc = ensure_tc_ws(h, nf);
    if (rc == CD_OK) rc = ws_acquire(h, s);
    if (rc != CD_OK) return rc;
    for (int64_t b0 = 0; b0 < B; b0 += slab) {
        const int64_t n = (B - b0) < slab ? (B - b0) : slab;
        float* feat = d_feat_opt ? d_feat_opt + b0 * felems
                                 : (tcp ? nullptr : reinterpret_cast<float*>(reinterpret_cast<char*>(h->d_ws) + cnn_ws));
        rc = i16 ? launch_features_i16(h, d_audio16 + b0 * stride_or_hop, stride_or_hop, mode, coef, n, feat,
                                       tcp ? cnn_tc_map_ptr(h) : nullptr, s)
                 : launch_features(h, d_audio + b0 * stride_or_hop, nullptr, nullptr, stride_or_hop, mode, coef, n, feat,
                                   tcp ? cnn_tc_map_ptr(h) : nullptr, s);
        if (rc != CD_OK) return rc;
        rc = tcp ? launch_cnn_tc(h, nullptr, n, d_logits + b0 * CD_N_CLASSES, s, nf)
                 : launch_cnn_fp32(h, feat, n, d_logits + b0 * CD_N_CLASSES, s, nf);
        if (rc != CD_OK) return rc;
    }
    return ws_release(h, s);
}

int cd_e2e_forward(cd_handle* h, const float* d_audio, int64_t stride, int64_t B, float* d_logits, float* d_feat_opt,
                   int flags, float pre_emphasis_coef, int precision, void* stream) {
    if (!h) return CD_E_ARG;
    if (B < 0 || stride < kSeg || (B > 0 && (!d_audio || !d_logits)))
        return set_err(h, CD_E_ARG, "cd_e2e_forward: bad argument");
    if (precision != CD_PRECISION_FP32 && precision != CD_PRECISION_BF16)
        return set_err(h, CD_E_ARG, "precision must be CD_PRECISION_FP32 or CD_PRECISION_BF16");
    if (!feat_flags_ok(h, flags, pre_emphasis_coef, "cd_e2e_forward")) return CD_E_ARG;
    DeviceGuard g(h->device);
    return e2e_impl(h, d_audio, false, stride, kModeFixed | feat_mode_bits(flags), pre_emphasis_coef, B, d_logits, d_feat_opt,
                    precision, (cudaStream_t)stream);
}

int cd_e2e_forward_i16(cd_handle* h, const int16_t* d_audio, int64_t stride, int64_t B, float* d_logits, float* d_feat_opt,
                       int flags, float pre_emphasis_coef, int precision, void* stream) {
    if (!h) return CD_E_ARG;
    if (B < 0 || stride < kSeg || (stride & 7) || (B > 0 && (!d_audio || !d_logits)))
        return set_err(h, CD_E_ARG, "cd_e2e_forward_i16: bad argument (need stride >= 16000 and a multiple of 8)");
    if (precision != CD_PRECISION_FP32 && precision != CD_PRECISION_BF16)
        return set_err(h, CD_E_ARG, "precision must be CD_PRECISION_FP32 or CD_PRECISION_BF16");
    if (!feat_flags_ok(h, flags, pre_emphasis_coef, "cd_e2e_forward_i16")) return CD_E_ARG;
    DeviceGuard g(h->device);
    return e2e_impl(h, d_audio, true, stride, kModeFixed | feat_mode_bits(flags), pre_emphasis_coef, B, d_logits, d_feat_opt,
                    precision, (cudaStream_t)stream);
}

int cd_e2e_forward_stream(cd_handle* h, const float* d_audio, int64_t n_samples, int64_t hop, int64_t n_windows,
                          float* d_logits, int flags, float pre_emphasis_coef, int precision, void* stream) {
    if (!h) return CD_E_ARG;
    if (hop <= 0 || n_windows != cd_stream_windows(n_samples, hop) || (n_windows > 0 && (!d_audio || !d_logits)))
        return set_err(h, CD_E_ARG, "cd_e2e_forward_stream: bad argument");
    if (precision != CD_PRECISION_FP32 && precision != CD_PRECISION_BF16)
        return set_err(h, CD_E_ARG, "precision must be CD_PRECISION_FP32 or CD_PRECISION_BF16");
    if (!feat_flags_ok(h, flags, pre_emphasis_coef, "cd_e2e_forward_stream")) return CD_E_ARG;
    DeviceGuard g(h->device);
    return e2e_impl(h, d_audio, false, hop, kModeStream | feat_mode_bits(flags | CD_FEAT_NORMALIZE), pre_emphasis_coef, n_windows,
                    d_logits, nullptr, precision, (cudaStream_t)stream);
}

// Host buffers in, host logits out.  Three-deep pipeline: chunk i is copied host->device on stream i%3 while
// chunk i-1 computes and chunk i-2's logits drain.  All three streams share the handle's workspace, so
// compute of consecutive chunks is serialised with events; copies overlap.
static int e2e_host_impl(cd_handle* h, const void* h_audio, bool i16, int64_t B, float* h_logits, int flags, float coef,
                         int precision);

int cd_e2e_forward_host(cd_handle* h, const float* h_audio, int64_t B, float* h_logits, int flags,
                        float pre_emphasis_coef, int precision) {
    return e2e_host_impl(h, h_audio, false, B, h_logits, flags, pre_emphasis_coef, precision);
}

int cd_e2e_forward_host_i16(cd_handle* h, const int16_t* h_audio, int64_t B, float* h_logits, int flags,
                            float pre_emphasis_coef, int precision) {
    return e2e_host_impl(h, h_audio, true, B, h_logits, flags, pre_emphasis_coef, precision);
}

static int e2e_host_impl(cd_handle* h, const void* h_audio, bool i16, int64_t B, float* h_logits, int flags, float coef,
                         int precision) {
    if (!h) return CD_E_ARG;
    if (B < 0 || (B > 0 && (!h_audio || !h_logits))) return set_err(h, CD_E_ARG, "cd_e2e_forward_host: bad argument");
    if (!h->weights_loaded) return set_err(h, CD_E_STATE, "cd_e2e_forward_host: call cd_cnn_load first");
    if (precision != CD_PRECISION_FP32 && precision != CD_PRECISION_BF16)
        return set_err(h, CD_E_ARG, "precision must be CD_PRECISION_FP32 or CD_PRECISION_BF16");
    if (!feat_flags_ok(h, flags, coef, "cd_e2e_forward_host")) return CD_E_ARG;
    const int mode = kModeFixed | feat_mode_bits(flags);
    const int nf = (mode & kModeDeltaDelta) ? kFeatDD : kFeat;
    if (B == 0) return CD_OK;
    DeviceGuard g(h->device);
    // copy/compute overlap granularity: a chunk's compute (~0.15 ms) hides under its own copy (1.2 ms of fp32), and only
    // the last chunk's compute is exposed -- measured 51.99 / 53.18 / 53.95 / 54.05 / 54.21 GB/s at 8192 / 4096 / 2048 / 1024 / 512
    const int64_t chunk = 1024;
    if (!h->pipe_streams[0]) {
        for (int i = 0; i < 3; ++i) {
            CD_CUDA(h, cudaStreamCreateWithFlags(&h->pipe_streams[i], cudaStreamNonBlocking));
            CD_CUDA(h, cudaEventCreateWithFlags(&h->pipe_events[i], cudaEventDisableTiming));
            CD_CUDA(h, cudaMalloc(&h->dev_in[i], (size_t)chunk * kSeg * sizeof(float)));
        }
    }
    if (h->pipe_logits_cap < B) {
        CD_CUDA(h, cudaDeviceSynchronize());
        cudaFree(h->dev_logits);
        h->dev_logits = nullptr;
        h->pipe_logits_cap = 0;
        CD_CUDA(h, cudaMalloc(&h->dev_logits, (size_t)B * CD_N_CLASSES * sizeof(float)));
        h->pipe_logits_cap = B;
    }
    const bool tcp = precision == CD_PRECISION_BF16;
    const size_t cnn_ws = tcp ? cnn_tc_ws_bytes(chunk) : cnn_fp32_ws_bytes(chunk);
    int rc = ensure_ws(h, cnn_ws + (tcp ? 0 : (size_t)chunk * nf * kFrames * sizeof(float)));
    if (rc == CD_OK && tcp) rc = ensure_tc_ws(h, nf);
    if (rc != CD_OK) return rc;
    float* feat = tcp ? nullptr : reinterpret_cast<float*>(reinterpret_cast<char*>(h->d_ws) + cnn_ws);
    int k = 0;
    cudaEvent_t prev_compute = nullptr;
    for (int64_t b0 = 0; b0 < B; b0 += chunk, k = (k + 1) % 3) {
        const int64_t n = (B - b0) < chunk ? (B - b0) : chunk;
        cudaStream_t s = h->pipe_streams[k];
        const size_t esz = i16 ? sizeof(int16_t) : sizeof(float);
        CD_CUDA(h, cudaMemcpyAsync(h->dev_in[k], static_cast<const char*>(h_audio) + (size_t)b0 * kSeg * esz, (size_t)n * kSeg * esz,
                                   cudaMemcpyHostToDevice, s));
        if (prev_compute) CD_CUDA(h, cudaStreamWaitEvent(s, prev_compute, 0));   // shared workspace
        else if ((rc = ws_acquire(h, s)) != CD_OK) return rc;                     // ... whoever used it before this call
        rc = i16 ? launch_features_i16(h, reinterpret_cast<const int16_t*>(h->dev_in[k]), kSeg, mode, coef, n, feat,
                                       tcp ? cnn_tc_map_ptr(h) : nullptr, s)
                 : launch_features(h, h->dev_in[k], nullptr, nullptr, kSeg, mode, coef, n, feat,
                                   tcp ? cnn_tc_map_ptr(h) : nullptr, s);
        if (rc != CD_OK) return rc;
        rc = tcp ? launch_cnn_tc(h, nullptr, n, h->dev_logits + b0 * CD_N_CLASSES, s, nf)
                 : launch_cnn_fp32(h, feat, n, h->dev_logits + b0 * CD_N_CLASSES, s, nf);
        if (rc != CD_OK) return rc;
        CD_CUDA(h, cudaEventRecord(h->pipe_events[k], s));
        prev_compute = h->pipe_events[k];
        CD_CUDA(h, cudaMemcpyAsync(h_logits + b0 * CD_N_CLASSES, h->dev_logits + b0 * CD_N_CLASSES,
                                   (size_t)n * CD_N_CLASSES * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < 3; ++i) CD_CUDA(h, cudaStreamSynchronize(h->pipe_streams[i]));
    h->ws_busy = false;                                      // everything this call enqueued has completed
    return CD_OK;
}

}  // extern "C"
