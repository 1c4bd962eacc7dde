// Minimal C++ caller of the C ABI (include/cough_b200.h): audio on the device -> logits, no Python, no torch.
//   make examples && ./build/e2e_minimal            (needs a B200; weights are random, the point is the call sequence)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cough_b200.h"

#define CK(call)                                                                        \
    do {                                                                                \
        int rc_ = (call);                                                               \
        if (rc_ != 0) {                                                                 \
            std::fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, cd_last_error(h)); \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

static std::vector<float> rnd(size_t n, float scale, unsigned seed) {
    std::vector<float> v(n);
    for (size_t i = 0; i < n; ++i) {
        seed = seed * 1664525u + 1013904223u;
        v[i] = scale * ((seed >> 8) * (1.0f / 8388608.0f) - 1.0f);
    }
    return v;
}

int main() {
    cd_handle* h = nullptr;
    if (cd_create(0, &h) != 0) {
        std::fprintf(stderr, "cd_create: %s\n", cd_last_error(nullptr));
        return 1;
    }
    // weights in the reference's state_dict layout (src/model.py:210-247): here random, BatchNorm = identity
    struct Layer {
        std::vector<float> w, b, g, beta, mean, var;
        Layer(int cout, int cin_kk, float scale, unsigned seed)
            : w(rnd((size_t)cout * cin_kk, scale, seed)), b(rnd(cout, 0.1f, seed + 100)), g(cout, 1.f), beta(cout, 0.f), mean(cout, 0.f), var(cout, 1.f) {}
        cd_conv_bn view() const { return {w.data(), b.data(), g.data(), beta.data(), mean.data(), var.data()}; }
    };
    Layer stem(32, 49, 0.1f, 1), c01(64, 32 * 9, 0.05f, 2), c02(64, 64 * 9, 0.05f, 3), s0(64, 32, 0.1f, 4);
    Layer c11(128, 64 * 9, 0.05f, 5), c12(128, 128 * 9, 0.03f, 6), s1(128, 64, 0.1f, 7);
    auto wfc = rnd(2 * 128, 0.5f, 8), bfc = rnd(2, 0.1f, 9);
    cd_residual_weights w{};
    w.stem = stem.view();
    w.b0_conv1 = c01.view();
    w.b0_conv2 = c02.view();
    w.b0_skip = s0.view();
    w.b1_conv1 = c11.view();
    w.b1_conv2 = c12.view();
    w.b1_skip = s1.view();
    w.fc_weight = wfc.data();
    w.fc_bias = bfc.data();
    CK(cd_cnn_load(h, &w));

    const int64_t B = 256;
    auto audio = rnd((size_t)B * CD_SEGMENT, 0.3f, 99);
    float *d_audio = nullptr, *d_logits = nullptr;
    cudaMalloc(&d_audio, audio.size() * sizeof(float));
    cudaMalloc(&d_logits, B * CD_N_CLASSES * sizeof(float));
    cudaMemcpy(d_audio, audio.data(), audio.size() * sizeof(float), cudaMemcpyHostToDevice);
    CK(cd_e2e_forward(h, d_audio, CD_SEGMENT, B, d_logits, nullptr, CD_FEAT_NORMALIZE, 0.f, CD_PRECISION_BF16, nullptr));
    std::vector<float> logits(B * CD_N_CLASSES);
    cudaMemcpy(logits.data(), d_logits, logits.size() * sizeof(float), cudaMemcpyDeviceToHost);
    std::printf("%s: clip 0 logits %.4f %.4f, %lld kernel launches\n", cd_version(), logits[0], logits[1],
                (long long)cd_launch_count(h));
    // the same from host memory, copies pipelined inside the library
    std::vector<float> host_logits(B * CD_N_CLASSES);
    CK(cd_e2e_forward_host(h, audio.data(), B, host_logits.data(), CD_FEAT_NORMALIZE, 0.f, CD_PRECISION_BF16));
    std::printf("host path agrees: %s\n", host_logits == logits ? "bit for bit" : "NO");
    cudaFree(d_audio);
    cudaFree(d_logits);
    cd_destroy(h);
    return 0;
}
