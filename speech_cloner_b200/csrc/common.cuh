// Shared declarations for the speech-cloner DSP kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scdsp {

constexpr int kHop = 80;               // fast-path hop (hop_length_ms 5.0 @ 16 kHz, hp/*.json)
constexpr int kFeThreads = 320;        // 16 units x 20 threads
constexpr int kFeUnits = 16;
constexpr int kFeFrames = 32;          // frames per front-end tile (one round)
constexpr int kFeSpan = kHop * (kFeFrames - 1) + 400;   // 2880 staged samples per tile
constexpr int kMaxMelChunks = 10;   // 16 frames x 10 chunks = all 160 compute threads of a pass-A tile
constexpr int kMaxMels = 128;

// per-utterance record, device resident
struct UttStat {
    float gain;          // mean_abs_amp_norm / mean|y|                      audio_lib.py:126
    unsigned p_max;      // float bits of max / min |X|^2 over the utterance  audio_lib.py:157, :231
    unsigned p_min;
    unsigned m_max;      // float bits of max / min mel power                 audio_lib.py:172, :235
    unsigned m_min;
    float pad[3];
};

// ragged-batch descriptor arrays (device), all of length n_utts (+1 for the prefixes)
struct Ragged {
    const int64_t* sample_off;   // first sample of utterance u in the packed waveform buffer
    const int64_t* sample_len;   // samples of utterance u
    const int64_t* frame_off;    // first row of utterance u in the packed feature buffers
    const int32_t* frame_cnt;    // frames of utterance u
    const int32_t* tile_prefix;  // exclusive prefix sum of tiles per utterance (n_utts + 1)
    int32_t n_utts;
    // pass A splits an utterance's tiles into interior tiles [int_first, int_first + int_count), whose
    // samples need no reflect padding (persistent cp.async kernel), and edge tiles (plain kernel)
    const int32_t* int_first;
    const int32_t* int_count;
};

// ---- asynchronous global -> shared copies (LDGSTS) and named barriers
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
template <int ID, int COUNT> __device__ __forceinline__ void bar_sync() {
    if (ID == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}

// largest u with prefix[u] <= tile
__device__ __forceinline__ int find_utt(const int32_t* __restrict__ prefix, int n, int tile) {
    int lo = 0, hi = n;           // invariant: prefix[lo] <= tile < prefix[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

// np.pad(.., mode='reflect') index for any integer position, n >= 1
__device__ __forceinline__ int64_t reflect_idx(int64_t q, int64_t n) {
    if (n == 1) return 0;
    const int64_t period = 2 * (n - 1);
    int64_t m = q % period;
    if (m < 0) m += period;
    return m >= n ? period - m : m;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 10*log10(x) for x > 0 through the MUFU log2 unit: abs error ~2e-5 dB, three orders below the
// 1e-3 dB that the 1e-5 output tolerance allows after the 0.01 scale (audio_lib.py:231).
// Where the raw power dB (audio_lib.py:157 before the top_db clip) is computed: 1 = pass A stores |X|^2 and pass B takes
// the logarithm while it clips and shifts (pass B is HBM-bound and has issue slots to spare, pass A has not);
// 0 = pass A stores dB.  Same operations on the same values either way: results are bit-identical.
#ifndef SC_DB_IN_PASS_B
#define SC_DB_IN_PASS_B 1
#endif
// Every caller clamps x to >= 1e-10, so the denormal pre-scaling that __log2f() adds around the MUFU is dropped.
__device__ __forceinline__ float db10(float x) {
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    return 3.0102999566398120f * l;
}
// value pass A stores for a power p / value pass B reads back as raw dB
__device__ __forceinline__ float pdb_store(float p) { return SC_DB_IN_PASS_B ? p : db10(fmaxf(p, 1e-10f)); }
__device__ __forceinline__ float pdb_load(float v) { return SC_DB_IN_PASS_B ? db10(fmaxf(v, 1e-10f)) : v; }

}  // namespace scdsp
