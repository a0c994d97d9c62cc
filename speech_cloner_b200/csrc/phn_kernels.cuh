// Frame-label alignment of the dataset-cache builders: calc_PHN_target (audio_lib.py:51-85), called once per
// utterance right after calc_MFCC_input in TIMIT_reader.py:192, ARCTIC_reader.py:157, TARGET_spk_reader.py:173.
// Integer interval arithmetic, one thread per frame: frame t looks at samples [t*hop - win/2, t*hop - win/2 + win)
// (the centred STFT window), finds the first phoneme interval whose end lies beyond the window start and keeps it
// unless the NEXT interval overlaps the window by strictly more samples.  The reference advances a cursor
// (`while phn_v[i][1] <= i_win_s`); with interval ends in non-decreasing order (what every .PHN / alignment file
// holds, and what the host checks) the cursor equals a binary search.
#pragma once
#include "common.cuh"

namespace scdsp {

struct PhnBatch {
    const int32_t* start;        // packed interval starts (samples)
    const int32_t* end;          // packed interval ends (samples, exclusive)
    const int64_t* phn_off;      // first interval of utterance u (n_utts + 1)
    const int64_t* frame_off;    // first output row of utterance u
    const int32_t* frame_cnt;    // frames of utterance u: 1 + len / hop
    int32_t n_utts, hop, win;
};

__global__ void __launch_bounds__(256) k_phn_target(PhnBatch pb, int32_t u0, int32_t* __restrict__ out) {
    const int u = u0 + blockIdx.y;
    if (u >= pb.n_utts) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= pb.frame_cnt[u]) return;
    const int64_t p0 = pb.phn_off[u];
    const int n = (int)(pb.phn_off[u + 1] - p0);
    const int32_t* __restrict__ s = pb.start + p0;
    const int32_t* __restrict__ e = pb.end + p0;
    const int64_t lo = (int64_t)t * pb.hop - pb.win / 2;
    const int64_t hi = lo + pb.win;
    // first c with e[c] > lo, clamped to the last interval
    int a = 0, b = n - 1;
    while (a < b) {
        const int mid = (a + b) >> 1;
        if ((int64_t)__ldg(e + mid) <= lo) a = mid + 1; else b = mid;
    }
    const int c = a;
    int pick = c;
    if (c + 1 < n) {
        const int64_t ov_a = min((int64_t)__ldg(e + c), hi) - max((int64_t)__ldg(s + c), lo);
        const int64_t ov_b = min((int64_t)__ldg(e + c + 1), hi) - max((int64_t)__ldg(s + c + 1), lo);
        if (ov_a < ov_b) pick = c + 1;
    }
    out[pb.frame_off[u] + t] = pick;
}

}  // namespace scdsp
