// SpecAugment on cached features: the device side of SpecAugment.__call__
// (src/augmentation.py:303-331), i.e. n_freq_masks x T.FrequencyMasking + n_time_masks x T.TimeMasking with mask
// value 0 (torchaudio functional.py mask_along_axis: rows / columns [start, end) are filled).  The intervals are sampled
// on the host in the reference's random-number order (cough_detector_b200/augmentation.py); this kernel only writes the
// zeros -- masked cells are stored, nothing is read, so a batch costs (masked fraction) x 36 KB per clip of HBM writes.
#include "cd_common.cuh"

namespace cd {

constexpr int kMaxMasks = 8;      // per axis

// masks: [B][2 * (nf + nt)] int32, per clip {start, end} of the nf row intervals, then of the nt column intervals;
// an empty interval (end <= start) masks nothing.  One CTA per clip.
__global__ void __launch_bounds__(256)
spec_augment_kernel(float* __restrict__ feat, int n_rows, int n_cols, int64_t B, const int32_t* __restrict__ masks, int nf, int nt) {
    const int64_t b = blockIdx.x;
    if (b >= B) return;
    __shared__ int32_t m[4 * kMaxMasks];
    const int nm = 2 * (nf + nt);
    if ((int)threadIdx.x < nm) m[threadIdx.x] = masks[b * nm + threadIdx.x];
    __syncthreads();
    float* f = feat + b * (int64_t)n_rows * n_cols;
    // row intervals: whole rows
    for (int k = 0; k < nf; ++k) {
        const int r0 = max(m[2 * k], 0), r1 = min(m[2 * k + 1], n_rows);
        for (int i = r0 * n_cols + threadIdx.x; i < r1 * n_cols; i += blockDim.x) f[i] = 0.f;
    }
    // column intervals: a strip of every row
    for (int k = 0; k < nt; ++k) {
        const int c0 = max(m[2 * (nf + k)], 0), c1 = min(m[2 * (nf + k) + 1], n_cols), w = c1 - c0;
        if (w <= 0) continue;
        for (int i = threadIdx.x; i < n_rows * w; i += blockDim.x) {
            const int r = i / w;
            f[r * n_cols + c0 + (i - r * w)] = 0.f;
        }
    }
}

int launch_spec_augment(cd_handle* h, float* d_feat, int n_rows, int n_cols, int64_t B, const int32_t* d_masks, int nf, int nt,
                        cudaStream_t s) {
    if (nf < 0 || nt < 0 || nf > kMaxMasks || nt > kMaxMasks) return set_err(h, CD_E_ARG, "cd_spec_augment: at most 8 masks per axis");
    if (B == 0 || nf + nt == 0) return CD_OK;
    spec_augment_kernel<<<(unsigned)B, 256, 0, s>>>(d_feat, n_rows, n_cols, B, d_masks, nf, nt);
    h->launches += 1;
    CD_CUDA(h, cudaGetLastError());
    return CD_OK;
}

}  // namespace cd
