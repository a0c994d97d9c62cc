// Window samplers of the training readers on a feature cache that is resident in HBM:
// Sound_DS.spec_window_sampler (sound_ds.py:262-350) and TIMIT.window_sampler (TIMIT_reader.py:474-523) cut one window
// of n_timesteps frames out of every sampled utterance - `ds_h5py[group][i_sample][i_s:i_e]` per feature group - and
// zero-pad utterances that are shorter (sound_ds.py:246-259, :303-311).  The random numbers (np.random.shuffle /
// np.random.randint on NumPy's global generator) stay on the host in the reference's order; what moves here is the data:
// one launch copies the windows of a batch out of every packed [rows][width] buffer (mfcc, mel_dB, power_dB, phn) into
// dense [batch][n_timesteps][width] tensors.  Pure HBM copy: a window is one contiguous run of valid*width words in the
// source and n_timesteps*width words in the destination, so threads walk it in 16-byte destination words.
#pragma once
#include "common.cuh"

namespace scdsp {

constexpr int kGatherMaxArrays = 4;
constexpr int kGatherThreads = 256;
constexpr int kGatherWordsPerBlock = 4 * 4 * kGatherThreads;   // four 16-byte words per thread

struct GatherArray {
    const uint32_t* src;     // packed [rows][width] buffer of 32-bit elements (float32 features, int32 labels)
    uint32_t* dst;           // dense [n_windows][n_timesteps][width]
    int64_t width;           // elements per row
};
struct GatherArgs {
    GatherArray a[kGatherMaxArrays];
};

// grid (blocks per window, windows of this launch from w0 on, n_arrays); first_row[w] = packed row of the window's first
// frame, valid[w] = rows that exist (< n_timesteps only for zero-padded utterances)
__global__ void __launch_bounds__(kGatherThreads)
k_window_gather(const __grid_constant__ GatherArgs g, const int64_t* __restrict__ first_row, const int32_t* __restrict__ valid, int32_t w0,
                int32_t n_timesteps, int64_t n_rows_total) {
    const GatherArray& A = g.a[blockIdx.z];
    const int w = w0 + blockIdx.y;
    const int64_t n_words = (int64_t)n_timesteps * A.width;
    int64_t r0 = first_row[w], v = valid[w];
    v = v < 0 ? 0 : (v > n_timesteps ? n_timesteps : v);
    if (r0 < 0 || r0 > n_rows_total) { r0 = 0; v = 0; }           // a window outside the cache reads nothing (zeros)
    if (r0 + v > n_rows_total) v = n_rows_total - r0;
    const int64_t n_valid = v * A.width;
    const uint32_t* __restrict__ s = A.src + r0 * A.width;
    uint32_t* __restrict__ d = A.dst + (int64_t)w * n_words;
    const int64_t stride = (int64_t)gridDim.x * kGatherThreads;
    const int64_t t0 = (int64_t)blockIdx.x * kGatherThreads + threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(d) & 15) != 0) {            // destination window not 16-byte aligned: word by word
        for (int64_t e = t0; e < n_words; e += stride) d[e] = e < n_valid ? __ldg(s + e) : 0u;
        return;
    }
    const bool src16 = (reinterpret_cast<uintptr_t>(s) & 15) == 0;
    const int64_t n4 = n_words >> 2;
    for (int64_t i = t0; i < n4; i += stride) {
        const int64_t e = 4 * i;
        uint4 q;
        if (e + 4 <= n_valid) {
            if (src16) {
                q = __ldg(reinterpret_cast<const uint4*>(s + e));
            } else {                                             // the window starts on any row: 4-byte loads, the warp
                q.x = __ldg(s + e); q.y = __ldg(s + e + 1);      // still covers 512 contiguous bytes
                q.z = __ldg(s + e + 2); q.w = __ldg(s + e + 3);
            }
        } else {
            q.x = e < n_valid ? __ldg(s + e) : 0u;
            q.y = e + 1 < n_valid ? __ldg(s + e + 1) : 0u;
            q.z = e + 2 < n_valid ? __ldg(s + e + 2) : 0u;
            q.w = 0u;                                            // e + 3 >= n_valid in this branch
        }
        *reinterpret_cast<uint4*>(d + e) = q;
    }
    if (blockIdx.x == 0) {
        const int64_t e = 4 * n4 + threadIdx.x;
        if (e < n_words) d[e] = e < n_valid ? __ldg(s + e) : 0u;
    }
}

}  // namespace scdsp
