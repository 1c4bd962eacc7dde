/* cough_b200.h -- C ABI of the B200-native cough_detector hot path (libcough_b200.so).
 *
 * The reference has no FFI: its hot path sits behind a Python class/function surface
 * (SURVEY.md §8b).  These entry points are what a thin binding for that surface calls; each
 * one names the reference interface it replaces (paths are into the reference repository).
 * Plain pointers and sizes only -- no torch types.  All device work is enqueued on the
 * caller's stream (a cudaStream_t passed as void*; NULL = legacy default stream) and is
 * asynchronous unless stated.  Every function returns 0 on success or a negative CD_E_* code;
 * cd_last_error() gives the message.  There is no CPU fallback: without a usable sm_100 device
 * cd_create fails.
 *
 * Ownership: pointers are borrowed for the duration of the call (device pointers: until the
 * enqueued work completes).  Inputs are never written.  A handle owns its constant tables,
 * folded weights (fp32 and bf16), resampling filters and workspaces; there is no per-device or
 * process-wide state, so calls on different handles -- also two handles on one GPU -- are
 * independent.  Calls on one handle are not thread-safe; they may use different streams (the
 * library orders successive users of the shared workspace with an event).  The only state a
 * handle carries between calls is its weights (cd_cnn_load): feature options travel with each call.
 */
#ifndef COUGH_B200_H
#define COUGH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CD_SAMPLE_RATE   16000
#define CD_SEGMENT       16000   /* AudioPreprocessor.segment_samples, src/preprocessing.py:82 */
#define CD_N_FEATURES    90      /* get_num_features(), src/preprocessing.py:536-550 (train.py:264-287 flags) */
#define CD_N_FRAMES      101     /* get_expected_time_frames(), src/preprocessing.py:532-534 */
#define CD_FEAT_ELEMS    (CD_N_FEATURES * CD_N_FRAMES)
#define CD_N_CLASSES     2

#define CD_OK            0
#define CD_E_ARG        -1       /* bad argument (NULL pointer, negative size, unsupported value) */
#define CD_E_CUDA       -2       /* a CUDA runtime call or kernel launch failed */
#define CD_E_NODEVICE   -3       /* no CUDA device / device is not sm_100 */
#define CD_E_STATE      -4       /* call order (e.g. CNN forward before weights were loaded) */
#define CD_E_NOMEM      -5

#define CD_FEAT_NORMALIZE    1   /* peak-normalise each clip before the features */
#define CD_FEAT_PRE_EMPHASIS 2   /* use_pre_emphasis: y[n] = x[n] - coef x[n-1] ahead of both feature blocks (:213-239, :452) */
#define CD_FEAT_PCEN         4   /* use_pcen: rows 0..63 are apply_pcen + per-clip min/max (:305-340, :400-404) instead of dB */
#define CD_FEAT_DELTA_DELTA  8   /* use_delta_delta: 13 more rows, compute_deltas(compute_deltas(mfcc)) (:463-466): d_feat is
                                    B x 103 x 101; the network then runs on 103 rows (cd_cnn_forward_rows / cd_e2e_*) */
#define CD_N_FEATURES_DD 103

#define CD_PRECISION_FP32 0      /* direct convolution on CUDA cores, fp32 weights + activations */
#define CD_PRECISION_BF16 1      /* implicit-GEMM tcgen05 tiles, bf16 operands, fp32 accumulate */

typedef struct cd_handle cd_handle;

/* One convolution + its BatchNorm, exactly as stored in the reference state_dict
 * (src/model.py:227-229, 274-283).  Host pointers, fp32, PyTorch layouts:
 * weight [cout][cin][k][k], everything else [cout]. */
typedef struct cd_conv_bn {
    const float* weight;
    const float* bias;
    const float* bn_weight;        /* gamma */
    const float* bn_bias;          /* beta */
    const float* bn_running_mean;
    const float* bn_running_var;
} cd_conv_bn;

/* The 51-tensor state_dict of CoughDetectorResidual (src/model.py:210-247) minus the
 * num_batches_tracked counters. */
typedef struct cd_residual_weights {
    cd_conv_bn stem;               /* conv1.0 / conv1.1          7x7 s2 p3, 1 -> 32   */
    cd_conv_bn b0_conv1;           /* res_blocks.0.conv1 / bn1   3x3 s2 p1, 32 -> 64  */
    cd_conv_bn b0_conv2;           /* res_blocks.0.conv2 / bn2   3x3 s1 p1, 64 -> 64  */
    cd_conv_bn b0_skip;            /* res_blocks.0.skip.0 / .1   1x1 s2,    32 -> 64  */
    cd_conv_bn b1_conv1;           /* res_blocks.1.conv1 / bn1   3x3 s2 p1, 64 -> 128 */
    cd_conv_bn b1_conv2;           /* res_blocks.1.conv2 / bn2   3x3 s1 p1, 128 -> 128*/
    cd_conv_bn b1_skip;            /* res_blocks.1.skip.0 / .1   1x1 s2,    64 -> 128 */
    const float* fc_weight;        /* fc.2.weight [2][128] */
    const float* fc_bias;          /* fc.2.bias   [2]      */
} cd_residual_weights;

/* A convolution / a BatchNorm as stored in a state_dict (host pointers, fp32, PyTorch layouts), for the two model types
 * below whose BatchNorm does not directly follow every convolution. */
typedef struct cd_conv {
    const float* weight;           /* [cout][cin / groups][k][k] */
    const float* bias;             /* [cout] */
} cd_conv;
typedef struct cd_bn {
    const float* weight;           /* gamma */
    const float* bias;             /* beta */
    const float* running_mean;
    const float* running_var;
} cd_bn;

/* CoughDetectorSmall (src/model.py:146-207; model_type 'small', the reference's default): features.0/.1, then three
 * depthwise-separable blocks features.{4,5,6} / {9,10,11} / {14,15,16} (depthwise 3x3, pointwise 1x1, BatchNorm),
 * classifier.1 (Linear 128 -> 64) and classifier.4 (Linear 64 -> 2). */
typedef struct cd_small_weights {
    cd_conv conv1;                 /* features.0   3x3 p1, 1 -> 16 */
    cd_bn bn1;                     /* features.1 */
    struct {
        cd_conv dw;                /* depthwise 3x3 p1, groups = channels */
        cd_conv pw;                /* pointwise 1x1 */
        cd_bn bn;
    } block[3];                    /* 16 -> 32, 32 -> 64, 64 -> 128 */
    const float* fc1_weight;       /* classifier.1.weight [64][128] */
    const float* fc1_bias;
    const float* fc2_weight;       /* classifier.4.weight [2][64] */
    const float* fc2_bias;
} cd_small_weights;

/* CoughDetector (src/model.py:43-143; model_type 'standard'): conv_layers.{0..3}.conv / .bn (3x3 p1, 1 -> 32 -> 64 -> 128
 * -> 256, each followed by ReLU and MaxPool2d(2)), fc.0 (Linear 256 -> 128), fc.3 (Linear 128 -> 2). */
typedef struct cd_standard_weights {
    cd_conv conv[4];
    cd_bn bn[4];
    const float* fc1_weight;       /* fc.0.weight [128][256] */
    const float* fc1_bias;
    const float* fc2_weight;       /* fc.3.weight [2][128] */
    const float* fc2_bias;
} cd_standard_weights;

/* ---- lifetime ------------------------------------------------------------------------- */

/* Create a handle on CUDA device `device`.  Replaces AudioPreprocessor.__init__ (transform /
 * table construction, src/preprocessing.py:94-127): uploads the Hann, mel-filterbank, DCT and
 * FFT twiddle tables. */
int cd_create(int device, cd_handle** out);
int cd_destroy(cd_handle* h);
/* Message for the last non-zero return on this handle (or for a failed cd_create when h==NULL). */
const char* cd_last_error(const cd_handle* h);
/* Library / build identification, e.g. "cough_b200 0.1 sm_100a". */
const char* cd_version(void);

/* Host-side copy of the constant tables the kernels use (no GPU needed; h may be NULL):
 * hann[400], fb[257*64] (dense, row = FFT bin), dct[64*13] (row = mel band).  Any pointer may be
 * NULL.  Lets a binding check them against torchaudio's. */
int cd_get_tables(float* hann, float* fb, float* dct);

/* ---- features: AudioPreprocessor ------------------------------------------------------- */

/* Feature options.  Every feature-computing call carries its own `flags` (CD_FEAT_* bits) and `pre_emphasis_coef`;
 * nothing is remembered on the handle, so two AudioPreprocessor objects with different flags (src/preprocessing.py:32-51:
 * the flags belong to each object) can share one handle and interleave calls freely.  flags = CD_FEAT_NORMALIZE with
 * coef ignored is the training configuration (src/train.py:264-287).  CD_FEAT_PRE_EMPHASIS / CD_FEAT_PCEN keep 90 rows;
 * with CD_FEAT_DELTA_DELTA a call writes 103 rows per clip and the cd_e2e_* calls run the network on 103 rows.
 * pre_emphasis_coef must be in [0, 1] when CD_FEAT_PRE_EMPHASIS is set.  use_spectral_contrast is not built (its rows
 * are NaN in the reference itself). */

/* B fixed-length clips.  d_audio: device, B rows of 16000 fp32 samples, row pitch `stride`
 * samples (>= 16000).  d_feat: device, B x 90 (103) x 101 fp32, row-major (feature, time).
 * CD_FEAT_NORMALIZE applies normalize() (src/preprocessing.py:199-212) first, as
 * dataset.__getitem__ / process() / add_audio() do; without it this is bare extract_features().
 * Per clip: [normalize] -> extract_features (:432-489):
 * mel dB [0,1] rows 0..63 (:387-412), normalised MFCC rows 64..76 (:414-430), deltas rows
 * 77..89 (:342-356). */
int cd_feat_forward(cd_handle* h, const float* d_audio, int64_t stride, int64_t B,
                    float* d_feat, int flags, float pre_emphasis_coef, void* stream);

/* Same as cd_feat_forward for int16 PCM clips (the sample format of wav files; torchaudio.load divides by 32768,
 * which is exact in fp32 and cancels in the peak normalisation, so with CD_FEAT_NORMALIZE the features are
 * bit-identical to the fp32 path on x / 32768).  stride in samples, a multiple of 8.  Additive: the reference has
 * no int16 entry point; it halves the bytes that have to reach the GPU. */
int cd_feat_forward_i16(cd_handle* h, const int16_t* d_audio, int64_t stride, int64_t B,
                        float* d_feat, int flags, float pre_emphasis_coef, void* stream);

/* B clips of arbitrary length described by (offset, length) into one device buffer:
 * clip i = d_audio[d_offsets[i] .. d_offsets[i]+d_lengths[i]).  Per clip: normalize over the
 * WHOLE clip -> pad_or_trim to 16000 (centre trim / symmetric zero pad, :358-385) ->
 * extract_features, i.e. AudioPreprocessor.process (:491-517) at 16 kHz mono (CD_FEAT_NORMALIZE is implied).
 * d_offsets / d_lengths are device arrays.  length 0 is allowed (all-zero clip). */
int cd_feat_forward_var(cd_handle* h, const float* d_audio, const int64_t* d_offsets,
                        const int32_t* d_lengths, int64_t B, float* d_feat, int flags,
                        float pre_emphasis_coef, void* stream);

/* Sliding windows over one stream: window j = d_audio[j*hop .. j*hop+16000), per-window peak
 * normalisation (CD_FEAT_NORMALIZE is implied), no pad/trim: RealtimePreprocessor.add_audio (:582-612).
 * n_windows must be cd_stream_windows(n_samples, hop).  d_feat: n_windows x 90 (103) x 101. */
int cd_feat_forward_stream(cd_handle* h, const float* d_audio, int64_t n_samples, int64_t hop,
                           int64_t n_windows, float* d_feat, int flags, float pre_emphasis_coef,
                           void* stream);
/* Number of windows add_audio emits for n_samples buffered samples (0 if n_samples < 16000). */
int64_t cd_stream_windows(int64_t n_samples, int64_t hop);

/* ---- resampling: AudioPreprocessor.resample ------------------------------------------------- */

/* Output length torchaudio's resample produces: ceil(new_sr * n_in / orig_sr) (rates reduced by their gcd). */
int64_t cd_resample_length(int64_t n_in, int orig_sr, int new_sr);

/* T.Resample(orig_sr, new_sr) with torchaudio's defaults (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99),
 * as src/preprocessing.py:146-183 applies it.  d_in: rows x n_in fp32 (e.g. one row per channel), d_out: rows x
 * cd_resample_length(n_in, ...).  The polyphase table of a rate pair is built on first use (synchronous upload). */
int cd_resample(cd_handle* h, const float* d_in, int64_t rows, int64_t n_in, int orig_sr, int new_sr,
                float* d_out, void* stream);

/* Batched resample() + to_mono() of B variable-length clips packed in ONE device buffer, one launch (the host side of
 * AudioPreprocessor.process for a list of files, src/preprocessing.py:506-510, src/dataset.py:151-152): clip i is
 * d_channels[i] rows of d_n_in[i] samples (row-major [C][n]) starting at element d_in_offsets[i] of d_in, all at
 * orig_sr; every row is resampled to new_sr like cd_resample, the rows are averaged, and the mono clip is written to
 * d_out + d_out_offsets[i] (cd_resample_length(d_n_in[i]) samples).  sample_type: 0 = fp32, 1 = int16 PCM (converted
 * on the fly; the scale cancels in the peak normalisation that follows).  orig_sr == new_sr is allowed (mono mix /
 * conversion only).  max_n_in >= every d_n_in[i] sizes the launch.  The descriptor arrays are device arrays.  The output
 * feeds cd_feat_forward_var directly. */
int cd_condition_var(cd_handle* h, const void* d_in, int sample_type, const int64_t* d_in_offsets,
                     const int32_t* d_channels, const int32_t* d_n_in, const int64_t* d_out_offsets, int64_t B,
                     int64_t max_n_in, int orig_sr, int new_sr, float* d_out, void* stream);

/* ---- model: CoughDetectorResidual ------------------------------------------------------ */

/* Fold BatchNorm (eval mode, eps 1e-5) into the convolutions, repack for the kernels and upload.
 * Replaces create_model('residual') + load_state_dict + .to(device) + .eval()
 * (src/model.py:296-316, src/inference.py:147-157).  Synchronous.  May be called again to swap
 * weights. */
int cd_cnn_load(cd_handle* h, const cd_residual_weights* w);

/* forward(): d_feat B x 1 x 90 x 101 fp32 -> d_logits B x 2 fp32 (src/model.py:253-259).
 * precision: CD_PRECISION_FP32 or CD_PRECISION_BF16. */
int cd_cnn_forward(cd_handle* h, const float* d_feat, int64_t B, float* d_logits,
                   int precision, void* stream);

/* The same forward() for an input of n_rows feature rows: 90, or 103 for a use_delta_delta feature set (the network is
 * fully convolutional up to its global average pool, src/model.py:236-259; the kernels are instantiated for these two
 * heights).  d_feat B x 1 x n_rows x 101 fp32. */
int cd_cnn_forward_rows(cd_handle* h, const float* d_feat, int n_rows, int64_t B, float* d_logits,
                        int precision, void* stream);

/* The reference's other two model types, fp32 on CUDA cores (off the north-star path; cnn_plain.cu): load = fold the
 * BatchNorms + upload (create_model('small' | 'standard') + load_state_dict + eval, src/model.py:296-316), forward =
 * d_feat B x 1 x n_rows x 101 fp32 -> d_logits B x 2 (src/model.py:198-201, 107-128).  Any n_rows >= 8 works (both
 * networks are fully convolutional up to their global average pool). */
int cd_small_load(cd_handle* h, const cd_small_weights* w);
int cd_small_forward(cd_handle* h, const float* d_feat, int n_rows, int64_t B, float* d_logits, void* stream);
int cd_standard_load(cd_handle* h, const cd_standard_weights* w);
int cd_standard_forward(cd_handle* h, const float* d_feat, int n_rows, int64_t B, float* d_logits, void* stream);

/* predict() on logits: softmax probabilities (B x 2), argmax class (B, int64) as in
 * src/model.py:261-265, and the batched decision P(cough) >= threshold (B, uint8) that stands in
 * for the per-window test of src/inference.py:229.  Any output may be NULL. */
int cd_predict(cd_handle* h, const float* d_logits, int64_t B, float threshold,
               float* d_probs, int64_t* d_preds, uint8_t* d_decisions, void* stream);

/* ---- training-side use of the features ----------------------- */

/* SpecAugment.__call__ (src/augmentation.py:303-331) on a batch of cached features, in place: per clip n_freq_masks row
 * intervals and n_time_masks column intervals are filled with 0, as T.FrequencyMasking / T.TimeMasking do
 * (torchaudio functional.py mask_along_axis, mask_value 0).  d_feat: B x n_rows x 101 fp32; d_masks: device,
 * B x 2 (n_freq_masks + n_time_masks) int32, per clip {start, end} of each row interval, then of each column interval
 * (end exclusive; end <= start masks nothing -- that is how a clip the reference leaves alone with probability 1 - p is
 * expressed).  The intervals are sampled by the caller in the reference's random-number order
 * (cough_detector_b200/augmentation.py); at most 8 masks per axis. */
int cd_spec_augment(cd_handle* h, float* d_feat, int n_rows, int64_t B, const int32_t* d_masks,
                    int n_freq_masks, int n_time_masks, void* stream);

/* ---- end to end ------------------------------------------------------------------------ */

/* audio -> logits in one call: cd_feat_forward + cd_cnn_forward without the features leaving
 * the GPU.  d_feat_opt may be NULL or a B x 90 (103) x 101 buffer that receives the features.
 * flags / pre_emphasis_coef as for cd_feat_forward. */
int cd_e2e_forward(cd_handle* h, const float* d_audio, int64_t stride, int64_t B,
                   float* d_logits, float* d_feat_opt, int flags, float pre_emphasis_coef,
                   int precision, void* stream);

/* cd_e2e_forward for int16 PCM clips on the device (stride in samples, a multiple of 8). */
int cd_e2e_forward_i16(cd_handle* h, const int16_t* d_audio, int64_t stride, int64_t B,
                       float* d_logits, float* d_feat_opt, int flags, float pre_emphasis_coef,
                       int precision, void* stream);

/* Same, over sliding windows of one stream (window j = d_audio[j*hop ..)); d_logits is
 * n_windows x 2. */
int cd_e2e_forward_stream(cd_handle* h, const float* d_audio, int64_t n_samples, int64_t hop,
                          int64_t n_windows, float* d_logits, int flags, float pre_emphasis_coef,
                          int precision, void* stream);

/* HOST buffers: h_audio B x 16000 fp32 (pageable or pinned), h_logits B x 2 fp32.  The library
 * pipelines chunked host->device copies, the kernels and the device->host copy of the logits on
 * its own streams, and returns when h_logits is complete (synchronous).  This is the call the
 * bench's end-to-end number times. */
int cd_e2e_forward_host(cd_handle* h, const float* h_audio, int64_t B, float* h_logits,
                        int flags, float pre_emphasis_coef, int precision);

/* cd_e2e_forward_host for int16 PCM host buffers (B x 16000 int16): half the host->device bytes. */
int cd_e2e_forward_host_i16(cd_handle* h, const int16_t* h_audio, int64_t B, float* h_logits,
                            int flags, float pre_emphasis_coef, int precision);

/* Page-locked host memory for the host-buffer calls above (callers without their own allocator).  flags:
 * CD_HOST_PINNED (cudaHostAllocDefault) or CD_HOST_WRITE_COMBINED (cudaHostAllocWriteCombined: faster for the device to
 * read on some hosts, slow for the CPU to read back -- use it for audio going in, never for logits coming out).
 * tools/h2d_matrix.py measures both on the box at hand. */
#define CD_HOST_PINNED         0
#define CD_HOST_WRITE_COMBINED 1
int cd_host_alloc(void** out, size_t bytes, int flags);
int cd_host_free(void* p);

/* Diagnostic: D[i][n] = sum_k A[i + shift][k] * B[n][k] (i < 128) on one tcgen05 tile through the same
 * shared-memory / instruction descriptors the convolution kernels use (operands rounded to bf16, fp32
 * accumulate).  `shift` starts the A operand that many 128-byte rows into the swizzled tile -- the
 * "shifted view" the convolutions use for their taps.  mode 0: SWIZZLE_128B rows, plain address offset (what the
 * 3x3 kernels do); mode 1: same plus the descriptor's base_offset field (kept as the negative control: it is NOT what
 * the hardware wants and gives wrong results for shifts that are not multiples of 8); mode 2: SWIZZLE_32B rows (the
 * stem's 4x4-pixel blocks); mode 3: one cta_group::2 MMA stream over a CTA pair (A has 256 rows, shift must be 0).
 * Host pointers; A has 128 + shift rows; N multiple of 32 in [32, 256], K multiple of 64, shift in [0, 64].
 * Synchronous. */
int cd_selftest_umma(cd_handle* h, int N, int K, int shift, int mode, const float* h_A, const float* h_B,
                     float* h_D);

/* Number of kernels this handle has launched since creation (the bench's gpu_launches). */
int64_t cd_launch_count(const cd_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* COUGH_B200_H */
