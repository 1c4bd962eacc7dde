// Shared declarations of libcough_b200 (internal; the public surface is include/cough_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/cough_b200.h"

namespace cd {

constexpr int kSeg      = CD_SEGMENT;    // 16000 samples per clip
constexpr int kFrames   = CD_N_FRAMES;   // 101
constexpr int kPairs    = 51;            // frame pairs packed into one complex FFT
constexpr int kMels     = 64;
constexpr int kMfcc     = 13;
constexpr int kFeat     = CD_N_FEATURES; // 90
constexpr int kFeatElem = CD_FEAT_ELEMS; // 9090
constexpr int kFeatDD     = CD_N_FEATURES_DD;   // 103 rows with use_delta_delta
constexpr int kFeatElemDD = kFeatDD * 101;
constexpr int kNfft     = 512;
constexpr int kWin      = 400;
constexpr int kHop      = 160;
// bf16 copy of the features as the stem consumes it ("blocked map"): the zero-padded image (feature rows
// -3..92, time cols -3..104) cut into 4x4 blocks; block (a, b) is one 32-byte row (16 bf16, k = 4 r + c) of a
// SWIZZLE_32B K-major operand, row index a * kBmapB + b, so the stem's nine 4x4-block taps are shifted views.
constexpr int kBmapA     = 24;                    // block rows  (map rows 0..95  = feature rows -3..92)
constexpr int kBmapB     = 27;                    // block cols  (map cols 0..107 = time cols   -3..104)
constexpr int kBmapRows  = kBmapA * kBmapB;       // 648 operand rows per clip
constexpr int kMapElems  = kBmapRows * 16;        // 10 368 bf16 = 20 736 B per clip
// element offset of map pixel (R, C) (R = feature row + 3, C = time col + 3) inside a clip's blocked map
__host__ __device__ inline int bmap_offset(int R, int C) {
    const int row = (R >> 2) * kBmapB + (C >> 2), r = R & 3, c = C & 3;
    return row * 16 + ((((r >> 1) ^ ((row >> 2) & 1)) << 3) | ((r & 1) * 4 + c));   // 16-byte chunk ^= row bit 2
}

// mel contraction as executed by a lane.  Every FFT bin feeds at most two adjacent triangular bands (m, m + 1); the
// bins are cut into 64 segments, segment m = the bins whose LOWER band is m (the falling side of band m = the rising
// side of band m + 1; <= 2 bins for m < 32, <= 4 above), and lane l owns segments l and l + 32.  It reads each of its
// bins once and forms own = sum W[m][k] P[k] and up = sum W[m+1][k] P[k]; band m = own(m) + up(m - 1), the second term
// arriving by one shuffle from the lane below -- 124 bin reads per frame instead of the filterbank's 244 taps.
constexpr int kMelSlotsLo = 2;
constexpr int kMelSlotsHi = 4;
constexpr int kMelSlots   = kMelSlotsLo + kMelSlotsHi;

struct FeatTables {                 // lives in device global memory; every lane keeps its slice in registers
    float2 twl[32][16];             // lane m: W512^(m r), r = 0..15 (stage-1 twiddles; staged in shared memory per CTA)
    float  win[kWin];               // 0.5 * hann (the 0.5 makes |Z +- conj Z'|^2 the power directly)
    float2 melt[kMelSlots][32];     // {W[m][k], W[m+1][k]} of the lane's s-th bin (slots 0-1: segment l, 2-5: segment l+32)
    int    melb[2][32];             // first bin of the lane's lo / hi segment | (bins in it) << 8
};

// Spatial shapes of the network for NF feature rows (101 frames): 90 (the training configuration and its
// pre-emphasis / PCEN variants) or 103 (use_delta_delta).  The width chain 101 -> 25 -> 13 -> 7 is fixed.
template <int NF>
struct NetShape {
    static_assert(NF == kFeat || NF == kFeatDD, "90 or 103 feature rows");
    static constexpr int H0 = ((NF - 1) / 2 + 1) / 2;     // stem: 7x7/s2 pad 3, then 2x2 max-pool (floor)   22 | 26
    static constexpr int H1 = (H0 - 1) / 2 + 1;           // block 0 (3x3/s2 pad 1)                          11 | 13
    static constexpr int H2 = (H1 - 1) / 2 + 1;           // block 1                                          6 | 7
    static constexpr int W0 = 25, W1 = 13, W2 = 7;
    static constexpr int BA = H0 + 2;                     // 4x4-pixel block rows of the stem's map          24 | 28
    static constexpr int BROWS = (BA * kBmapB + 7) & ~7;  // operand rows per clip, a multiple of 8 so that the SWIZZLE_32B
                                                          // pattern (row bit 2) is the same for every clip      648 | 760
    static constexpr int MAP_ELEMS = BROWS * 16;
    static constexpr int FEAT_ELEMS = NF * kFrames;
};
static_assert(NetShape<kFeat>::BA == kBmapA && NetShape<kFeat>::MAP_ELEMS == kMapElems, "90-row shapes");

// Folded network, fp32 (CUDA-core path).  Layouts chosen for the direct-conv kernels:
// weights [k*k*cin][cout] (cout fastest) so a warp reads them as broadcast float4s.
struct CnnWeightsF32 {
    float* stem_w;  float* stem_b;      // [49][32], [32]
    float* b0c1_w;  float* b0c1_b;      // [9*32][64]
    float* b0sk_w;  float* b0sk_b;      // [32][64]
    float* b0c2_w;  float* b0c2_b;      // [9*64][64]
    float* b1c1_w;  float* b1c1_b;      // [9*64][128]
    float* b1sk_w;  float* b1sk_b;      // [64][128]
    float* b1c2_w;  float* b1c2_b;      // [9*128][128]
    float* fc_w;    float* fc_b;        // [2][128], [2]
};

// Folded network as the tensor-core kernels read it: device pointers into cd_handle::d_wblob_bf16 (packed bf16 operand
// images followed by an fp32 blob of biases and the classifier).  Owned by the handle, rewritten by cd_cnn_load.
struct TcWeights {
    __nv_bfloat16 *b0c1, *b0sk, *b0c2, *b1c1, *b1sk, *b1c2;
    __nv_bfloat16* stem_w;
    float *stem_b, *b0c1_b, *b0sk_b, *b0c2_b, *b1c1_b, *b1sk_b, *b1c2_b, *fc_w, *fc_b;
};

// Polyphase filter bank of one (orig, new) rate pair (resample.cu); device buffers owned by the handle.
struct ResampleTable {
    int orig, neu, width, support;     // reduced rates, torchaudio's `width`, taps kept per phase
    int* d_k0;                         // [neu] first kept tap (index into the padded input window)
    float* d_taps;                     // [neu][support]
};

}  // namespace cd

namespace cd { struct PlainNet; }

struct cd_handle {
    int device = 0;
    int sm_count = 0;
    std::string err;
    int64_t launches = 0;

    cd::FeatTables* d_tables = nullptr;
    std::vector<cd::ResampleTable> resample_tables;   // built on first use of a rate pair, freed by cd_destroy

    bool weights_loaded = false;
    float* d_wblob = nullptr;            // one allocation holding every folded fp32 tensor
    cd::CnnWeightsF32 w32{};
    __nv_bfloat16* d_wblob_bf16 = nullptr;
    cd::TcWeights tcw{};                 // pointers into d_wblob_bf16
    cd::PlainNet* plain[2] = {nullptr, nullptr};   // folded 'small' / 'standard' networks (cnn_plain.cu)
    unsigned char* d_ws_tc_fam[2] = {nullptr, nullptr};   // tensor-core path activations per shape family (90 | 103 rows):
                                                          // fixed layout, pads zeroed once at allocation
    unsigned char* d_ws_tc = nullptr;    // the family the current call uses (set by ensure_tc_ws)

    float* d_ws_stream = nullptr;        // shared-STFT streaming: mel power of the global frames + block maxima
    size_t ws_stream_bytes = 0;

    // workspace (grown on demand, never shrunk)
    float* d_ws = nullptr;
    size_t ws_bytes = 0;
    // The workspaces (d_ws, d_ws_tc, d_ws_stream) are shared by every call on this handle.  Each launch sequence that
    // touches them ends by recording ws_event on its stream; the next one, if it runs on a DIFFERENT stream, first
    // makes that stream wait for the event (ws_acquire / ws_release in api.cu), so callers may alternate streams.
    cudaEvent_t ws_event = nullptr;
    cudaStream_t ws_stream = nullptr;
    bool ws_busy = false;

    // host-pipeline state for cd_e2e_forward_host
    cudaStream_t pipe_streams[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t pipe_events[3] = {nullptr, nullptr, nullptr};
    float* dev_in[3] = {nullptr, nullptr, nullptr};
    float* dev_logits = nullptr;
    int64_t pipe_logits_cap = 0;
};

namespace cd {

int set_err(cd_handle* h, int code, const char* what, cudaError_t e = cudaSuccess);
int ensure_ws(cd_handle* h, size_t bytes);

#define CD_CUDA(h, call)                                                     \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) return cd::set_err((h), CD_E_CUDA, #call, e__); \
    } while (0)

// kernels' host launchers (each returns a CD_* code and bumps h->launches)
// mode: FeatMode bits, every option explicit per launch (nothing is remembered on the handle); coef: pre_emphasis_coef
int launch_features(cd_handle* h, const float* d_audio, const int64_t* d_offsets, const int32_t* d_lengths,
                    int64_t stride_or_hop, int mode, float coef, int64_t B, float* d_feat, __nv_bfloat16* d_map,
                    cudaStream_t s);
int launch_spec_augment(cd_handle* h, float* d_feat, int n_rows, int n_cols, int64_t B, const int32_t* d_masks, int nf, int nt,
                        cudaStream_t s);
int launch_cnn_fp32(cd_handle* h, const float* d_feat, int64_t B, float* d_logits, cudaStream_t s, int nf = kFeat);
size_t cnn_fp32_ws_bytes(int64_t B);

enum FeatMode { kModeFixed = 0, kModeVar = 1, kModeStream = 2, kModeMask = 3, kModeNoNormalize = 4,
                kModePreEmph = 8, kModePcen = 16, kModeDeltaDelta = 32 };

}  // namespace cd
