// Sample-rate conversion on the GPU: AudioPreprocessor.resample (src/preprocessing.py:168-183), i.e.
// torchaudio T.Resample with its defaults (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99;
// torchaudio/functional/functional.py:1305-1432).  The polyphase filter bank is built on the host with
// the same arithmetic (float64, the phase term -j/new rounded through float32 like torch's int64/int
// division) and stored compactly: for output phase j only the taps inside the +-6 zero-crossing window
// are non-zero (about 34 of the 2*width+orig taps torchaudio's dense conv1d multiplies through).
#include <math.h>
#include <vector>

#include "cd_common.cuh"

namespace cd {

// The filter banks live in the handle (cd_handle::resample_tables): no process-wide state, freed by cd_destroy.
void free_resample_tables(cd_handle* h) {
    for (ResampleTable& t : h->resample_tables) {
        cudaFree(t.d_k0);
        cudaFree(t.d_taps);
    }
    h->resample_tables.clear();
}

static int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }
int64_t resample_length(int64_t n_in, int orig_sr, int new_sr);

static int get_table(cd_handle* h, int orig_sr, int new_sr, ResampleTable* out) {
    const int g = gcd_i(orig_sr, new_sr), orig = orig_sr / g, neu = new_sr / g;
    for (const ResampleTable& t : h->resample_tables)
        if (t.orig == orig && t.neu == neu) {
            *out = t;
            return CD_OK;
        }
    const int lpw = 6;
    const double base = (double)(orig < neu ? orig : neu) * 0.99;
    const int width = (int)ceil(lpw * (double)orig / base);
    const int ntap = 2 * width + orig;
    std::vector<double> full((size_t)neu * ntap);
    std::vector<int> k0(neu), k1(neu);
    int support = 1;
    for (int j = 0; j < neu; ++j) {
        const double phase = (double)((float)(-j) / (float)neu);       // torch: int64 arange / int -> float32
        int lo = ntap, hi = -1;
        for (int k = 0; k < ntap; ++k) {
            double t = (phase + (double)(k - width) / (double)orig) * base;
            if (t < -lpw) t = -lpw;
            if (t > lpw) t = lpw;
            const double c = cos(t * M_PI / lpw / 2.0);
            const double window = c * c;
            const double tp = t * M_PI;
            const double v = (tp == 0.0 ? 1.0 : sin(tp) / tp) * window * (base / (double)orig);
            full[(size_t)j * ntap + k] = v;
            if (fabs(t) < lpw) {                                       // outside: window = cos(pi/2)^2 ~ 4e-33
                if (k < lo) lo = k;
                if (k > hi) hi = k;
            }
        }
        k0[j] = lo;
        k1[j] = hi;
        if (hi - lo + 1 > support) support = hi - lo + 1;
    }
    std::vector<float> taps((size_t)neu * support, 0.f);
    for (int j = 0; j < neu; ++j)
        for (int s = 0; s < support && k0[j] + s <= k1[j]; ++s) taps[(size_t)j * support + s] = (float)full[(size_t)j * ntap + k0[j] + s];
    ResampleTable t{orig, neu, width, support, nullptr, nullptr};
    CD_CUDA(h, cudaMalloc(&t.d_k0, neu * sizeof(int)));
    CD_CUDA(h, cudaMalloc(&t.d_taps, taps.size() * sizeof(float)));
    CD_CUDA(h, cudaMemcpy(t.d_k0, k0.data(), neu * sizeof(int), cudaMemcpyHostToDevice));
    CD_CUDA(h, cudaMemcpy(t.d_taps, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice));
    h->resample_tables.push_back(t);
    *out = t;
    return CD_OK;
}

// y[f * neu + j] = sum_s taps[j][s] * x[f * orig + k0[j] + s - width]   (x = 0 outside [0, n_in))
__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ x, int64_t n_in, int64_t n_out, int orig, int neu, int width, int support,
                const int* __restrict__ k0, const float* __restrict__ taps, float* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t row = blockIdx.y;
    if (i >= n_out) return;
    const int64_t f = i / neu;
    const int j = (int)(i - f * neu);
    const float* xr = x + row * n_in;
    const float* tp = taps + (size_t)j * support;
    const int64_t base = f * orig + k0[j] - width;
    float acc = 0.f;
    for (int s = 0; s < support; ++s) {
        const int64_t k = base + s;
        const float v = (k >= 0 && k < n_in) ? __ldg(xr + k) : 0.f;
        acc = fmaf(__ldg(tp + s), v, acc);
    }
    y[row * n_out + i] = acc;
}

// Batched conditioning of variable-length clips, one launch: clip b = `chans[b]` rows of `n_in[b]` samples at
// in + in_off[b] (row-major [C][n]); every row is resampled like resample_kernel, the rows are averaged (to_mono,
// src/preprocessing.py:185-197: waveform.mean(dim=0) AFTER the resampling, :508-510) and the mono clip goes to
// out + out_off[b].  orig == neu: plain (convert +) channel mean.  T = float or int16 PCM.
template <class T>
__global__ void __launch_bounds__(256)
condition_var_kernel(const T* __restrict__ in, const int64_t* __restrict__ in_off, const int32_t* __restrict__ chans,
                     const int32_t* __restrict__ n_in, const int64_t* __restrict__ out_off, int orig, int neu, int width,
                     int support, const int* __restrict__ k0, const float* __restrict__ taps, float* __restrict__ out) {
    const int64_t b = blockIdx.y;
    const int64_t n = n_in[b];
    const int64_t n_out = (neu * n + orig - 1) / orig;
    const int C = chans[b];
    const T* x = in + in_off[b];
    float* y = out + out_off[b];
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_out; i += (int64_t)gridDim.x * 256) {
        float sum = 0.f;
        if (orig == neu) {
            for (int c = 0; c < C; ++c) sum += (float)x[c * n + i];
        } else {
            const int64_t f = i / neu;
            const int j = (int)(i - f * neu);
            const float* tp = taps + (size_t)j * support;
            const int64_t base = f * orig + k0[j] - width;
            for (int c = 0; c < C; ++c) {
                const T* xr = x + c * n;
                float acc = 0.f;
                for (int s = 0; s < support; ++s) {
                    const int64_t k = base + s;
                    const float v = (k >= 0 && k < n) ? (float)xr[k] : 0.f;
                    acc = fmaf(__ldg(tp + s), v, acc);
                }
                sum += acc;
            }
        }
        y[i] = C == 1 ? sum : sum / (float)C;
    }
}

int launch_condition_var(cd_handle* h, const void* d_in, bool i16, const int64_t* d_in_off, const int32_t* d_chans,
                         const int32_t* d_n_in, const int64_t* d_out_off, int64_t B, int64_t max_n_in, int orig_sr, int new_sr,
                         float* d_out, cudaStream_t s) {
    if (B == 0) return CD_OK;
    ResampleTable t{1, 1, 0, 0, nullptr, nullptr};
    if (orig_sr != new_sr) {
        int rc = get_table(h, orig_sr, new_sr, &t);
        if (rc != CD_OK) return rc;
    }
    const int64_t max_out = resample_length(max_n_in, orig_sr, new_sr);
    int64_t gx = (max_out + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > 4096) gx = 4096;
    for (int64_t b0 = 0; b0 < B; b0 += 65535) {                       // gridDim.y limit
        const int64_t nb = B - b0 < 65535 ? B - b0 : 65535;
        dim3 grid((unsigned)gx, (unsigned)nb);
        if (i16)
            condition_var_kernel<int16_t><<<grid, 256, 0, s>>>(static_cast<const int16_t*>(d_in), d_in_off + b0, d_chans + b0,
                                                               d_n_in + b0, d_out_off + b0, t.orig, t.neu, t.width, t.support,
                                                               t.d_k0, t.d_taps, d_out);
        else
            condition_var_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(d_in), d_in_off + b0, d_chans + b0,
                                                             d_n_in + b0, d_out_off + b0, t.orig, t.neu, t.width, t.support,
                                                             t.d_k0, t.d_taps, d_out);
        h->launches += 1;
    }
    CD_CUDA(h, cudaGetLastError());
    return CD_OK;
}

int64_t resample_length(int64_t n_in, int orig_sr, int new_sr) {
    const int g = gcd_i(orig_sr, new_sr);
    const int64_t orig = orig_sr / g, neu = new_sr / g;
    return (neu * n_in + orig - 1) / orig;                            // ceil(new * length / orig)
}

int launch_resample(cd_handle* h, const float* d_in, int64_t rows, int64_t n_in, int orig_sr, int new_sr, float* d_out,
                    cudaStream_t s) {
    ResampleTable t;
    int rc = get_table(h, orig_sr, new_sr, &t);
    if (rc != CD_OK) return rc;
    const int64_t n_out = resample_length(n_in, orig_sr, new_sr);
    if (n_out == 0 || rows == 0) return CD_OK;
    dim3 grid((unsigned)((n_out + 255) / 256), (unsigned)rows);
    resample_kernel<<<grid, 256, 0, s>>>(d_in, n_in, n_out, t.orig, t.neu, t.width, t.support, t.d_k0, t.d_taps, d_out);
    h->launches += 1;
    CD_CUDA(h, cudaGetLastError());
    return CD_OK;
}

}  // namespace cd
