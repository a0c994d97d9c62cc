// Generic-size kernels: any even n_fft <= 4096, any hop_length, any win_length <= n_fft.
//
// The reference accepts arbitrary STFT geometry (the calc_MFCC_input defaults are hop 40 / win 400,
// from_power_to_wav's are hop 40 / win 800, audio_lib.py:92-93, :281-282) although every shipped
// hp/*.json uses n_fft = 400, hop = 80.  These kernels keep the whole API on the GPU for the other
// geometries with a direct O(N^2) real DFT per frame (exact table twiddles, index reduced mod N):
// far from the FFT-400 kernels' throughput, but still GPU-resident and never a CPU fallback.
#pragma once
#include "common.cuh"
#include "fe_kernels.cuh"
#include "gl_kernels.cuh"

namespace scdsp {

constexpr int kGenMaxNfft = 2048;
constexpr int kGenThreads = 256;
constexpr int kGenFeFrames = 4;     // frames per front-end tile
constexpr int kGenGlGroup = 2;      // frames transformed together in the Griffin-Lim kernel

struct GenTables {
    const double* win;      // analysis window, centre padded to n_fft (float64, like the reference's FFT input)
    const cxd* wn;          // exp(-2*pi*i*m/n_fft) in float64
    int32_t n_fft, n_bins, hop;
};

struct GenGlTables {
    const float* win;       // hann, centre padded to n_fft
    const cxf* wn;
    const double* win_sq;   // hann^2 (float64)
    const float* inv_wss;   // steady-state 1 / sum-square, period hop
    int32_t n_fft, n_bins, hop;
};

__host__ __device__ inline size_t gen_align(size_t x) { return (x + 15) & ~size_t(15); }

inline size_t gen_fe_smem_bytes(int n_fft, int hop, int n_mels) {
    (void)hop;
    const int bins = 1 + n_fft / 2;
    size_t s = 0;
    s += gen_align(sizeof(double) * kGenFeFrames * n_fft);         // windowed frames
    s += gen_align(sizeof(cxd) * n_fft);                           // twiddles
    s += gen_align(sizeof(float) * kGenFeFrames * bins);           // power
    s += gen_align(sizeof(float2) * bins);                         // mel weights
    s += gen_align(sizeof(int32_t) * (n_mels + 2));                // mel interval starts
    s += gen_align(sizeof(float) * 4 * (kGenThreads / 32));        // reduction scratch
    s += gen_align(sizeof(float) * kGenFeFrames * (n_mels + 1));   // mel dB
    return s;
}

// direct real DFT of `xw` (n_fft windowed samples) at bin k
template <typename R>
__device__ __forceinline__ cx<R> dft_bin(const R* __restrict__ xw, const cx<R>* __restrict__ wn, int n_fft, int k) {
    R re = 0, im = 0;
    int idx = 0;
    for (int n = 0; n < n_fft; ++n) {
        const cx<R> w = wn[idx];
        const R x = xw[n];
        re = sc_fma(x, w.x, re);
        im = sc_fma(x, w.y, im);
        idx += k;
        if (idx >= n_fft) idx -= n_fft;
    }
    return mk<R>(re, im);
}

__global__ void __launch_bounds__(kGenThreads)
k_gen_fe_pass_a(const float* __restrict__ wav, Ragged rg, GenTables gt, FeTables tb, FeParams prm,
                UttStat* __restrict__ stat, float* __restrict__ pdb_out, float* __restrict__ mel_raw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n_fft = gt.n_fft, bins = gt.n_bins, hop = gt.hop, n_mels = tb.n_mels;
    unsigned char* sp = smem_raw;
    double* xw = reinterpret_cast<double*>(sp);          sp += gen_align(sizeof(double) * kGenFeFrames * n_fft);
    cxd* wn = reinterpret_cast<cxd*>(sp);                sp += gen_align(sizeof(cxd) * n_fft);
    float* power = reinterpret_cast<float*>(sp);         sp += gen_align(sizeof(float) * kGenFeFrames * bins);
    float2* mel_w = reinterpret_cast<float2*>(sp);       sp += gen_align(sizeof(float2) * bins);
    int32_t* istart = reinterpret_cast<int32_t*>(sp);    sp += gen_align(sizeof(int32_t) * (n_mels + 2));
    float (*red)[kGenThreads / 32] = reinterpret_cast<float (*)[kGenThreads / 32]>(sp);
    sp += gen_align(sizeof(float) * 4 * (kGenThreads / 32));
    float* mel_db = reinterpret_cast<float*>(sp);

    const int tid = threadIdx.x;
    const int u = find_utt(rg.tile_prefix, rg.n_utts, blockIdx.x);
    const int t0 = (blockIdx.x - rg.tile_prefix[u]) * kGenFeFrames;
    const int T = rg.frame_cnt[u];
    const int nfr = min(kGenFeFrames, T - t0);
    const int64_t L = rg.sample_len[u];
    const float* __restrict__ y = wav + rg.sample_off[u];
    const float gain = stat[u].gain;
    const double c = prm.pre_emphasis;

    for (int e = tid; e < nfr * n_fft; e += kGenThreads) {
        const int f = e / n_fft, n = e - f * n_fft;
        const int64_t r = reflect_idx((int64_t)(t0 + f) * hop + n - n_fft / 2, L);
        const float cur = gain * __ldg(y + r);
        const float prev = r > 0 ? gain * __ldg(y + r - 1) : 0.0f;
        xw[e] = ((double)cur - c * (double)prev) * __ldg(gt.win + n);
    }
    for (int i = tid; i < n_fft; i += kGenThreads) wn[i] = gt.wn[i];
    for (int i = tid; i < bins; i += kGenThreads) mel_w[i] = tb.mel_w[i];
    for (int i = tid; i < n_mels + 2; i += kGenThreads) istart[i] = tb.mel_istart[i];
    __syncthreads();
    for (int e = tid; e < nfr * bins; e += kGenThreads) {
        const int f = e / bins, k = e - f * bins;
        const cxd x = dft_bin<double>(xw + f * n_fft, wn, n_fft, k);
        power[e] = (float)fma(x.x, x.x, x.y * x.y);
    }
    __syncthreads();
    fe_epilogue_a<kGenThreads>(power, kGenFeFrames, bins, nfr, mel_w, istart, tb, mel_db, red, stat + u,
                               pdb_out + (rg.frame_off[u] + t0) * bins, mel_raw + (rg.frame_off[u] + t0) * n_mels);
}

// ---------------------------------------------------------------------------------------------
// Generic Griffin-Lim iteration.  A tile owns `out_per_tile` consecutive output samples and walks
// all frames that touch them in ascending order, kGenGlGroup at a time; each frame's windowed
// inverse transform is added into a shared accumulator before the next one (ordered, atomics-free).
__host__ __device__ inline int gen_gl_out_per_tile(int n_fft, int hop) {
    const int ov = (n_fft + hop - 1) / hop;
    return hop * (ov > 8 ? ov : 8);
}
inline size_t gen_gl_smem_bytes(int n_fft, int hop) {
    const int bins = 1 + n_fft / 2;
    size_t s = 0;
    s += gen_align(sizeof(float) * kGenGlGroup * n_fft);                         // windowed frames / time output
    s += gen_align(sizeof(cxf) * kGenGlGroup * bins);                         // spectra
    s += gen_align(sizeof(cxf) * n_fft);                                      // twiddles
    s += gen_align(sizeof(float) * n_fft);                                       // window
    s += gen_align(sizeof(float) * (gen_gl_out_per_tile(n_fft, hop)));           // accumulator
    return s;
}

__device__ __forceinline__ float gen_inv_wss_at(int64_t p, int T, const GenGlTables& tb) {
    const int n_fft = tb.n_fft, hop = tb.hop;
    int64_t lo = p >= n_fft ? (p - n_fft) / hop + 1 : 0;
    int64_t hi = p / hop;
    const int64_t full_hi = hi;
    if (hi > T - 1) hi = T - 1;
    if (p >= n_fft - 1 && hi == full_hi) return __ldg(tb.inv_wss + (int)(p % hop));   // all covering frames exist
    float acc = 0.f;
    for (int64_t i = lo; i <= hi; ++i) acc = (float)((double)acc + __ldg(tb.win_sq + (int)(p - i * hop)));
    return acc > 1.1754944e-38f ? 1.0f / acc : 1.0f;
}

__global__ void __launch_bounds__(kGenThreads)
k_gen_gl_iter(const GlJob* __restrict__ jobs, int n_jobs, const int32_t* __restrict__ tile_prefix, GenGlTables tb,
              const float* __restrict__ amp, const float* __restrict__ phase0, const float* __restrict__ wav_in,
              float* __restrict__ wav_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n_fft = tb.n_fft, bins = tb.n_bins, hop = tb.hop, half = n_fft / 2;
    const int out_per_tile = gen_gl_out_per_tile(n_fft, hop);
    unsigned char* sp = smem_raw;
    float* xt = reinterpret_cast<float*>(sp);        sp += gen_align(sizeof(float) * kGenGlGroup * n_fft);
    cxf* spec = reinterpret_cast<cxf*>(sp);       sp += gen_align(sizeof(cxf) * kGenGlGroup * bins);
    cxf* wn = reinterpret_cast<cxf*>(sp);         sp += gen_align(sizeof(cxf) * n_fft);
    float* win = reinterpret_cast<float*>(sp);       sp += gen_align(sizeof(float) * n_fft);
    float* acc = reinterpret_cast<float*>(sp);

    const int tid = threadIdx.x;
    const int ji = find_utt(tile_prefix, n_jobs, blockIdx.x);
    const GlJob job = jobs[ji];
    const int T = job.T;
    const int64_t Lw = (int64_t)hop * (T - 1);
    const int64_t p_first = ((job.out_first + half) / out_per_tile) * out_per_tile;
    const int64_t o = p_first + (int64_t)(blockIdx.x - job.tile0) * out_per_tile;   // padded start of the tile
    const bool init = phase0 != nullptr;

    for (int i = tid; i < n_fft; i += kGenThreads) { wn[i] = tb.wn[i]; win[i] = tb.win[i]; }
    for (int i = tid; i < out_per_tile; i += kGenThreads) acc[i] = 0.f;
    // frames touching padded [o, o + out_per_tile)
    int64_t f_first = o >= n_fft ? (o - n_fft) / hop + 1 : 0;
    int64_t f_last = (o + out_per_tile - 1) / hop;
    if (f_last > T - 1) f_last = T - 1;
    const float* __restrict__ src = init ? nullptr : wav_in + job.wav_in_off;
    const float inv_n = 1.0f / (float)n_fft;
    __syncthreads();

    for (int64_t fg = f_first; fg <= f_last; fg += kGenGlGroup) {
        const int ng = (int)((f_last - fg + 1) < kGenGlGroup ? (f_last - fg + 1) : kGenGlGroup);
        if (!init) {
            for (int e = tid; e < ng * n_fft; e += kGenThreads) {
                const int g = e / n_fft, n = e - g * n_fft;
                const int64_t r = reflect_idx((fg + g) * hop + n - half, Lw) - job.wav_in_first;
                const float v = (r >= 0 && r < job.wav_in_count) ? __ldg(src + r) : 0.0f;
                xt[e] = v * win[n];
            }
            __syncthreads();
        }
        // spectrum with the target magnitude imposed (:268-270) or the initial state (:256)
        for (int e = tid; e < ng * bins; e += kGenThreads) {
            const int g = e / bins, k = e - g * bins;
            const int64_t f = fg + g;
            cxf s = mk<float>(0.f, 0.f);
            if (f >= job.f_lo && f < (int64_t)job.f_lo + job.f_cnt) {
                const int64_t row = job.amp_row0 + (f - job.f_lo);
                const float a = __ldg(amp + row * bins + k);
                if (init) {
                    s = polar(a, __ldg(phase0 + row * bins + k));
                } else {
                    s = impose(dft_bin<float>(xt + g * n_fft, wn, n_fft, k), a);
                }
                if (k == 0 || k == bins - 1) s.y = 0.f;      // istft keeps only the real part there
            }
            spec[e] = s;
        }
        __syncthreads();
        // inverse real DFT, window; x[n] = (1/N) (S0 + (-1)^n S_{N/2} + 2 sum_k Re(S_k e^{+i 2 pi k n / N}))
        for (int e = tid; e < ng * n_fft; e += kGenThreads) {
            const int g = e / n_fft, n = e - g * n_fft;
            const cxf* __restrict__ s = spec + g * bins;
            float a = 0.f;
            int idx = n;
            for (int k = 1; k < bins - 1; ++k) {
                const cxf w = wn[idx];
                a = fmaf(s[k].x, w.x, a);
                a = fmaf(s[k].y, w.y, a);
                idx += n;
                if (idx >= n_fft) idx -= n_fft;
            }
            const float nyq = (n & 1) ? -s[bins - 1].x : s[bins - 1].x;
            xt[e] = (s[0].x + nyq + 2.0f * a) * inv_n * win[n];
        }
        __syncthreads();
        // ordered overlap-add: one frame after the other
        for (int g = 0; g < ng; ++g) {
            const int64_t base = (fg + g) * hop - o;          // tile-local position of the frame start
            for (int n = tid; n < n_fft; n += kGenThreads) {
                const int64_t l = base + n;
                if (l >= 0 && l < out_per_tile) acc[l] += xt[g * n_fft + n];
            }
            __syncthreads();
        }
    }
    float* __restrict__ dst = wav_out + job.wav_out_off;
    const int64_t out_end = job.out_first + job.out_count;
    for (int i = tid; i < out_per_tile; i += kGenThreads) {
        const int64_t p = o + i;
        const int64_t s = p - half;
        if (s < job.out_first || s >= out_end || s >= Lw) continue;
        dst[s - job.out_first] = acc[i] * gen_inv_wss_at(p, T, tb);
    }
}

}  // namespace scdsp
