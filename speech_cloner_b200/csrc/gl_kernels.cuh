// Griffin-Lim kernels (fast path: n_fft = 400, hop = 80): griffin_lim_alg, audio_lib.py:249-274,
// and the prologue / epilogue of from_power_to_wav, audio_lib.py:278-308.
//
// State between iterations is the WAVEFORM (float32, hop*(T-1) samples), not the complex
// spectrogram: one launch of k_gl_iter does  wav -> reflect pad -> STFT -> A * X/|X| -> iSTFT ->
// overlap-add -> / window-sum-square -> wav'  for a tile of 28 hops.  Each CTA recomputes the 4
// halo frames it shares with its neighbours, so the overlap-add needs neither atomics nor a second
// pass; per sample the covering frames are summed in ascending frame order like librosa.istft's loop.
#pragma once
#include "common.cuh"
#include "fft400.cuh"

namespace scdsp {

#ifndef SC_GL_SPLIT_TASKS
#define SC_GL_SPLIT_TASKS 0      // 1: packed columns c = 0 on warp 0 and c = 10 on warp 1 (each next to one-frame columns)
#endif
constexpr int kGlFrames = 32;                          // frames per tile (16 units x 2)
constexpr int kGlOutHops = kGlFrames - 4;              // complete hops per tile
constexpr int kGlOut = kGlOutHops * kHop;              // 2240 output samples per tile
constexpr int kGlSpan = kHop * (kGlFrames - 1) + 400;  // 2880
constexpr int kGlSeg = 480;                            // samples written by one unit (2 frames)

// one signal (utterance, or one rank's chunk of a long-form signal)
struct GlJob {
    int64_t amp_row0;     // row of frame `f_lo` in the amp / phase0 buffers
    int64_t wav_in_off;   // element offset of this job's input waveform
    int64_t wav_in_first; // whole-signal sample index of input element 0
    int64_t wav_in_count; // input elements available
    int64_t wav_out_off;  // element offset of output sample `out_first`
    int64_t out_first;    // whole-signal index of the first sample to write
    int64_t out_count;    // samples to write
    int32_t f_lo, f_cnt;  // frames whose amp rows are present
    int32_t T;            // frames of the whole signal
    int32_t tile0;        // exclusive prefix of tiles
};

struct GlTables {
    const cxf* w400;
    const float* win_half;     // 0.5 * hann (forward), centre padded to 400
    const float* win_inv;      // hann / 400 (inverse)
    const double* win_sq;      // hann^2 in float64 (window_sumsquare terms)
    const float* inv_wss;      // steady-state 1 / sum-square, period = hop
    const float* zero_row;     // 201 zeros: magnitude row of a frame that does not exist
};

// 1 / window_sumsquare at padded position p (librosa 0.6 filters.window_sumsquare, float32
// accumulator receiving float64 terms in ascending frame order); 1 where the sum is <= tiny.
// Positions are 32-bit (signals up to 2^31 samples, checked on the host): 64-bit divisions cost
// ~100 instructions each on the GPU and used to dominate this kernel.
__device__ __forceinline__ float inv_wss_at(int p, int T, const GlTables& tb) {
    const int lo = p >= kNfft ? (p - kNfft) / kHop + 1 : 0;
    int hi = p / kHop;
    if (hi > T - 1) hi = T - 1;
    if (hi - lo == kNfft / kHop - 1) return __ldg(tb.inv_wss + (p % kHop));
    float acc = 0.f;
    for (int i = lo; i <= hi; ++i) acc = (float)((double)acc + __ldg(tb.win_sq + (p - i * kHop)));
    return acc > 1.1754944e-38f ? 1.0f / acc : 1.0f;
}

struct GlSmem {
    float span[kGlSpan];
    float win_half[kNfft];
    float win_inv[kNfft];
    cxf slots[kFeUnits * kUnitSlots];
    float seg[kFeUnits * kGlSeg];       // windowed inverse transforms of the units (A + B pre-added), 480 each
};

template <bool INIT>
__global__ void __launch_bounds__(kFeThreads, 2)
k_gl_iter(const GlJob* __restrict__ jobs, int n_jobs, const int32_t* __restrict__ tile_prefix, GlTables tb,
          const float* __restrict__ amp, const float* __restrict__ phase0, const float* __restrict__ wav_in,
          float* __restrict__ wav_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GlSmem& sm = *reinterpret_cast<GlSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int ji = find_utt(tile_prefix, n_jobs, blockIdx.x);
    const GlJob job = jobs[ji];
    const int T = job.T;
    const int Lw = kHop * (T - 1);                            // whole-signal length (< 2^31, host-checked)
    // tiles sit on a whole-signal grid of kGlOut padded samples (padded = trimmed + 200), so a
    // time-chunked run pairs and sums frames exactly like the unchunked one (bit-identical)
    const int out_first = (int)job.out_first, out_end = (int)(job.out_first + job.out_count);
    const int p_first = ((out_first + kNfft / 2) / kGlOut) * kGlOut;
    const int o = p_first + (int)(blockIdx.x - job.tile0) * kGlOut;
    const int t0 = o / kHop - 4;                               // first frame of the tile
    const int span0 = t0 * kHop;                               // padded position of span[0]

    for (int i = tid; i < kNfft; i += kFeThreads) {
        sm.win_half[i] = 2.0f * tb.win_half[i];       // full Hann window
        sm.win_inv[i] = tb.win_inv[i];
    }
    if (!INIT) {
        const float* __restrict__ src = wav_in + job.wav_in_off;
        const int q0 = span0 - kNfft / 2;                      // whole-signal index of span[0]
        const int in_first = (int)job.wav_in_first, in_end = (int)(job.wav_in_first + job.wav_in_count);
        if (q0 >= 0 && q0 + kGlSpan <= Lw && q0 >= in_first && q0 + kGlSpan <= in_end) {
            const float* __restrict__ s0 = src + (q0 - in_first);       // interior tile: plain copy
            for (int i = tid; i < kGlSpan; i += kFeThreads) sm.span[i] = __ldg(s0 + i);
        } else {
            for (int i = tid; i < kGlSpan; i += kFeThreads) {
                const int r = (int)reflect_idx((int64_t)q0 + i, Lw) - in_first;
                sm.span[i] = (r >= 0 && r < in_end - in_first) ? __ldg(src + r) : 0.0f;
            }
        }
    }
    const int unit = tid / kUnitThreads;
    const int j = tid - unit * kUnitThreads;
    Twiddle tw;
    load_twiddles(tw, tb.w400, j);
    cxf* unit_slots = sm.slots + unit * kUnitSlots;
    __syncthreads();

    // Frames of this unit.  A frame that does not exist (outside [0, T) or outside the rows this job
    // holds) is fed exact zeros end to end: the two frames of a pair share one packed transform, so
    // any garbage in the partner would change the rounding of the real frame and break the
    // bit-identity of time-chunked runs.
    const int fa = t0 + 2 * unit, fb = fa + 1;
    const bool va = fa >= job.f_lo && fa < job.f_lo + job.f_cnt && fa < T;
    const bool vb = fb >= job.f_lo && fb < job.f_lo + job.f_cnt && fb < T;
    const float ka = va ? 1.0f : 0.0f, kb = vb ? 1.0f : 0.0f;
    if (!INIT) {
        float s[24];
        const float* __restrict__ src = sm.span + unit * (2 * kHop) + j;
#pragma unroll
        for (int m = 0; m < 24; ++m) s[m] = src[20 * m];
        float xa[20], xb[20];
#pragma unroll
        for (int n1 = 0; n1 < 20; ++n1) {
            const float w = sm.win_half[20 * n1 + j];
            xa[n1] = s[n1] * (w * ka);
            xb[n1] = s[n1 + 4] * (w * kb);
        }
        fwd_step1_real(xa, xb, tw, unit_slots + j);
        __syncthreads();
    }
    {
        // step-2 task of this thread (packed columns on warp 0, see step2_task) and its unit's two frames
        int u2, c2;
        step2_task<false>(tid, kFeUnits, u2, c2);
        const int ga = t0 + 2 * u2, gb = ga + 1;
        const bool wa = ga >= job.f_lo && ga < job.f_lo + job.f_cnt && ga < T;
        const bool wb = gb >= job.f_lo && gb < job.f_lo + job.f_cnt && gb < T;
        const int64_t ra = job.amp_row0 + (wa ? ga - job.f_lo : 0);
        const int64_t rb = job.amp_row0 + (wb ? gb - job.f_lo : 0);
        const float* __restrict__ amp_a = wa ? amp + ra * kBins : tb.zero_row;
        const float* __restrict__ amp_b = wb ? amp + rb * kBins : tb.zero_row;
        cxf* row = sm.slots + u2 * kUnitSlots + c2 * kSlotLd;
        cxf v[20];
        if (INIT) {
            gl_init_state(v, c2, amp_a, amp_b, wa ? phase0 + ra * kBins : tb.zero_row,
                          wb ? phase0 + rb * kBins : tb.zero_row);
        } else {
            fwd_step2(v, row);
            gl_update(v, c2, amp_a, amp_b);
        }
        inv_step2(v, row);                          // a slot row is read and rewritten by the same thread only
    }
    __syncthreads();
    float comb[24];
    {
        float ya[20], yb[20];
        inv_step1_real(ya, yb, tw, unit_slots + j);
#pragma unroll
        for (int m = 0; m < 24; ++m) {
            float a = 0.f;
            if (m < 20) a = ya[m] * sm.win_inv[20 * m + j];
            if (m >= 4) a += yb[m - 4] * sm.win_inv[20 * (m - 4) + j];
            comb[m] = a;
        }
    }
    float* seg = sm.seg;
    {
        float* __restrict__ dst = seg + unit * kGlSeg + j;
#pragma unroll
        for (int m = 0; m < 24; ++m) dst[20 * m] = comb[m];
    }
    __syncthreads();
    // ---- ordered overlap-add gather + normalisation; local positions [320, 320 + kGlOut) are complete
    {
        float* __restrict__ dst = wav_out + job.wav_out_off;
        // every output sample of an interior tile is covered by all 5 frames: periodic 1 / sum-square
        const bool steady = t0 >= 0 && t0 + kGlFrames <= T;
        // same static form as k_gl_iter_persist (units uh-2, uh-1, uh at offsets r+320, r+160, r; ascending-unit sums)
        static_assert(kGlOut == 7 * kFeThreads && kFeThreads == 4 * kHop && kGlSeg == 6 * kHop, "tile geometry");
        const int h = tid >= 2 * kHop ? 1 : 0;
        const int r = tid - 2 * kHop * h;
        const float* __restrict__ sg = seg + h * kGlSeg + r;
        float v0[7], v1[7], v2[7];
#pragma unroll
        for (int it = 0; it < 7; ++it) {
            v0[it] = sg[(2 * it) * kGlSeg + 4 * kHop];
            v1[it] = sg[(2 * it + 1) * kGlSeg + 2 * kHop];
            v2[it] = sg[(2 * it + 2) * kGlSeg];
        }
        const float nrm_steady = __ldg(tb.inv_wss + (r % kHop));
#pragma unroll
        for (int it = 0; it < 7; ++it) {
            const int l = 320 + tid + kFeThreads * it;              // local padded offset in the tile
            const int p = span0 + l;                                // padded position
            const int s = p - kNfft / 2;                            // whole-signal sample index
            if (s < out_first || s >= out_end || s >= Lw) continue;
            float acc = 0.f;
            acc += v0[it];
            acc += v1[it];
            acc += v2[it];
            const float nrm = steady ? nrm_steady : inv_wss_at(p, T, tb);
            dst[s - out_first] = acc * nrm;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent form of the iteration kernel: a CTA loops over tiles of a host-built (job, tile) table, keeps
// the windows / twiddles resident and requests the NEXT tile's 2 880 waveform samples with cp.async while
// it computes the current one (edge tiles, which need reflect padding, are staged synchronously).  The
// per-tile arithmetic is the body of k_gl_iter<false>, so both kernels give bit-identical waveforms.
struct GlSmemP {
    alignas(16) float span[2][kGlSpan];
    float win_half[kNfft];
    float win_inv[kNfft];
    cxf slots[kFeUnits * kUnitSlots];
    float seg[kFeUnits * kGlSeg];
    GlJob job[2];                       // descriptor of the staged tile: read from shared memory by the next iteration
    int2 ent[2];
};

struct GlGeom {
    int T, Lw, out_first, out_end, t0, span0;
};
__device__ __forceinline__ GlGeom gl_geom(const GlJob& job, int tile_in_job) {
    GlGeom g;
    g.T = job.T;
    g.Lw = kHop * (job.T - 1);
    g.out_first = (int)job.out_first;
    g.out_end = (int)(job.out_first + job.out_count);
    const int p_first = ((g.out_first + kNfft / 2) / kGlOut) * kGlOut;
    const int o = p_first + tile_in_job * kGlOut;
    g.t0 = o / kHop - 4;
    g.span0 = g.t0 * kHop;
    return g;
}

__global__ void __launch_bounds__(kFeThreads, 2)
k_gl_iter_persist(const GlJob* __restrict__ jobs, const int2* __restrict__ tile_tab, int n_tiles, GlTables tb,
                  const float* __restrict__ amp, const float* __restrict__ wav_in, float* __restrict__ wav_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GlSmemP& sm = *reinterpret_cast<GlSmemP*>(smem_raw);
    const int tid = threadIdx.x;
    for (int i = tid; i < kNfft; i += kFeThreads) {
        sm.win_half[i] = 2.0f * tb.win_half[i];       // full Hann window
        sm.win_inv[i] = tb.win_inv[i];
    }
    const int unit = tid / kUnitThreads;
    const int j = tid - unit * kUnitThreads;
    Twiddle tw;
    load_twiddles(tw, tb.w400, j);
    cxf* unit_slots = sm.slots + unit * kUnitSlots;

    auto stage = [&](int tile, int b) {
        const int2 e = tile_tab ? __ldg(tile_tab + tile) : make_int2(0, tile);   // no table: one job, tile = index
        const GlJob jb = jobs[e.x];
        if (tid == 0) { sm.job[b] = jb; sm.ent[b] = e; }
        const GlGeom g = gl_geom(jb, e.y);
        const float* __restrict__ src = wav_in + jb.wav_in_off;
        const int q0 = g.span0 - kNfft / 2;
        const int in_first = (int)jb.wav_in_first, in_end = (int)(jb.wav_in_first + jb.wav_in_count);
        float* dst = sm.span[b];
        if (q0 >= 0 && q0 + kGlSpan <= g.Lw && q0 >= in_first && q0 + kGlSpan <= in_end) {
            const float* __restrict__ s0 = src + (q0 - in_first);
            if ((reinterpret_cast<uintptr_t>(s0) & 15) == 0) {
                for (int i = tid; i < kGlSpan / 4; i += kFeThreads) cp_async16(dst + 4 * i, s0 + 4 * i);
            } else {
                for (int i = tid; i < kGlSpan; i += kFeThreads) cp_async4(dst + i, s0 + i);
            }
        } else {
            for (int i = tid; i < kGlSpan; i += kFeThreads) {
                const int r = (int)reflect_idx((int64_t)q0 + i, g.Lw) - in_first;
                dst[i] = (r >= 0 && r < in_end - in_first) ? __ldg(src + r) : 0.0f;
            }
        }
    };

    // step 1 of a tile: windowed frame pair of this unit -> first 20-point transforms -> exchange slots.
    // Frames that do not exist (outside [0, T) or outside the rows this job holds) are fed exact zeros end to end: the
    // two frames of a pair share one packed transform, so any garbage in the partner would change the rounding of the
    // real frame and break the bit-identity of time-chunked runs.
    auto step1 = [&](const GlJob& job, const GlGeom& g, const float* __restrict__ span) {
        const int fa = g.t0 + 2 * unit, fb = fa + 1;
        const bool va = fa >= job.f_lo && fa < job.f_lo + job.f_cnt && fa < g.T;
        const bool vb = fb >= job.f_lo && fb < job.f_lo + job.f_cnt && fb < g.T;
        const float ka = va ? 1.0f : 0.0f, kb = vb ? 1.0f : 0.0f;
        float s[24];
        const float* __restrict__ src = span + unit * (2 * kHop) + j;
#pragma unroll
        for (int m = 0; m < 24; ++m) s[m] = src[20 * m];
        float xa[20], xb[20];
#pragma unroll
        for (int n1 = 0; n1 < 20; ++n1) {
            const float w = sm.win_half[20 * n1 + j];
            xa[n1] = s[n1] * (w * ka);
            xb[n1] = s[n1 + 4] * (w * kb);
        }
        fwd_step1_real(xa, xb, tw, unit_slots + j);
    };

    // Three CTA-wide barriers per tile: the overlap-add gather of tile i (shared loads + global stores) and step 1 of
    // tile i+1 (shared loads + FMAs) touch different buffers (seg / next span -> slots) and run as one phase.
    int tile = blockIdx.x, b = 0;
    if (tile < n_tiles) stage(tile, 0);
    cp_async_wait_all();
    __syncthreads();
    if (tile < n_tiles) {
        const GlJob job0 = sm.job[0];
        step1(job0, gl_geom(job0, sm.ent[0].y), sm.span[0]);
    }
    __syncthreads();
#pragma unroll 1
    for (; tile < n_tiles; tile += gridDim.x, b ^= 1) {
        const int2 e = sm.ent[b];                     // written by stage() one iteration (and at least one barrier) ago
        const GlJob job = sm.job[b];
        const bool has_next = tile + (int)gridDim.x < n_tiles;
        if (has_next) stage(tile + gridDim.x, b ^ 1);
        const GlGeom g = gl_geom(job, e.y);
        const int T = g.T, Lw = g.Lw, out_first = g.out_first, out_end = g.out_end, t0 = g.t0, span0 = g.span0;
        {
            // step-2 task of this thread (packed columns on warp 0, see step2_task) and its unit's two frames
            int u2, c2;
            step2_task<SC_GL_SPLIT_TASKS != 0>(tid, kFeUnits, u2, c2);
            const int ga = t0 + 2 * u2, gb = ga + 1;
            const bool wa = ga >= job.f_lo && ga < job.f_lo + job.f_cnt && ga < T;
            const bool wb = gb >= job.f_lo && gb < job.f_lo + job.f_cnt && gb < T;
            const int64_t ra = job.amp_row0 + (wa ? ga - job.f_lo : 0);
            const int64_t rb = job.amp_row0 + (wb ? gb - job.f_lo : 0);
            const float* __restrict__ amp_a = wa ? amp + ra * kBins : tb.zero_row;
            const float* __restrict__ amp_b = wb ? amp + rb * kBins : tb.zero_row;
            cxf* row = sm.slots + u2 * kUnitSlots + c2 * kSlotLd;
            cxf v[20];
            fwd_step2(v, row);
            gl_update(v, c2, amp_a, amp_b);
            inv_step2(v, row);                          // a slot row is read and rewritten by the same thread only
        }
        __syncthreads();
        float* seg = sm.seg;
        {
            float comb[24];
            float ya[20], yb[20];
            inv_step1_real(ya, yb, tw, unit_slots + j);
#pragma unroll
            for (int m = 0; m < 24; ++m) {
                float a = 0.f;
                if (m < 20) a = ya[m] * sm.win_inv[20 * m + j];
                if (m >= 4) a += yb[m - 4] * sm.win_inv[20 * (m - 4) + j];
                comb[m] = a;
            }
            float* __restrict__ dst = seg + unit * kGlSeg + j;
#pragma unroll
            for (int m = 0; m < 24; ++m) dst[20 * m] = comb[m];
        }
        cp_async_wait_all();                            // this thread's part of the next span has landed ...
        __syncthreads();                                // ... and everybody's is visible; seg complete; slots free
        // ---- ordered overlap-add gather + normalisation; local positions [320, 320 + kGlOut) are complete
        {
            float* __restrict__ dst = wav_out + job.wav_out_off;
            // every output sample of an interior tile is covered by all 5 frames: periodic 1 / sum-square
            const bool steady = t0 >= 0 && t0 + kGlFrames <= T;
            // Output i = tid + 320 * it (it = 0..6) sits at local offset l = 320 + i = 160 * uh + r with
            // uh = 2 + 2 * it + (tid >= 160), r = tid mod 160: exactly the units uh-2, uh-1, uh (all in 0..15) cover it,
            // at segment offsets r + 320, r + 160, r.  Same ascending-unit summation as the generic gather of
            // k_gl_iter, with the index arithmetic folded away and all 21 loads issued first.
            static_assert(kGlOut == 7 * kFeThreads && kFeThreads == 4 * kHop && kGlSeg == 6 * kHop, "tile geometry");
            const int h = tid >= 2 * kHop ? 1 : 0;
            const int r = tid - 2 * kHop * h;
            const float* __restrict__ sg = seg + h * kGlSeg + r;
            float v0[7], v1[7], v2[7];
#pragma unroll
            for (int it = 0; it < 7; ++it) {
                v0[it] = sg[(2 * it) * kGlSeg + 4 * kHop];
                v1[it] = sg[(2 * it + 1) * kGlSeg + 2 * kHop];
                v2[it] = sg[(2 * it + 2) * kGlSeg];
            }
            const float nrm_steady = __ldg(tb.inv_wss + (r % kHop));
#pragma unroll
            for (int it = 0; it < 7; ++it) {
                const int l = 320 + tid + kFeThreads * it;              // local padded offset in the tile
                const int p = span0 + l;                                // padded position
                const int s = p - kNfft / 2;                            // whole-signal sample index
                if (s < out_first || s >= out_end || s >= Lw) continue;
                float acc = 0.f;
                acc += v0[it];
                acc += v1[it];
                acc += v2[it];
                const float nrm = steady ? nrm_steady : inv_wss_at(p, T, tb);
                dst[s - out_first] = acc * nrm;
            }
        }
        if (has_next) {
            const GlJob jn = sm.job[b ^ 1];
            step1(jn, gl_geom(jn, sm.ent[b ^ 1].y), sm.span[b ^ 1]);
        }
        __syncthreads();
    }
}

// sqrt(mean((a - b)^2)) per job: the value the reference prints when verbose (:262-264)
__global__ void __launch_bounds__(256) k_rms_delta_partial(const float* __restrict__ a, const float* __restrict__ b,
                                                          const GlJob* __restrict__ jobs, int n_jobs,
                                                          const int32_t* __restrict__ prefix,
                                                          double* __restrict__ partial) {
    const int ji = find_utt(prefix, n_jobs, blockIdx.x);
    const GlJob job = jobs[ji];
    const int64_t begin = (int64_t)(blockIdx.x - prefix[ji]) * 8192;
    const int64_t end = begin + 8192 < job.out_count ? begin + 8192 : job.out_count;
    const float* __restrict__ pa = a + job.wav_out_off;
    const float* __restrict__ pb = b + job.wav_out_off;
    double s = 0.0;
    for (int64_t i = begin + threadIdx.x; i < end; i += 256) {
        const float d = pa[i] - pb[i];
        s += (double)(d * d);
    }
    s = warp_sum(s);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        partial[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(128) k_rms_delta_final(const GlJob* __restrict__ jobs, int n_jobs,
                                                        const int32_t* __restrict__ prefix,
                                                        const double* __restrict__ partial, float* __restrict__ out,
                                                        int n_iters, int iter) {
    const int ji = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (ji >= n_jobs) return;
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    for (int t = prefix[ji] + lane; t < prefix[ji + 1]; t += 32) s += partial[t];
    s = warp_sum(s);
    if (lane == 0) out[(int64_t)ji * n_iters + iter] = (float)sqrt(s / (double)jobs[ji].out_count);
}

// ---------------------------------------------------------------------------------------------
// from_power_to_wav prologue (:290-298).
// The two means of the `realse` power law (:293, :296) are summed in a CANONICAL order that does not depend on how a
// long spectrogram is cut into per-rank chunks: one float64 partial per block of `block_rows` frames on the
// whole-signal grid (fixed thread / warp order inside a block), and a sequential sum over the blocks in order.
// A time-chunked run (ranks all_gather their block partials) therefore gets bit-identical scale factors.
struct P2aJob {
    int64_t row0;       // first row of this job in the p / amp buffers
    int64_t rows;       // frames present
    int32_t tile0;      // prefix of apply tiles (kP2aChunk elements each)
    int32_t blk0;       // prefix of sum blocks
};
constexpr int kP2aChunk = 16384;   // elements per apply tile

// partial sums of max(0,P) and max(0,P)^realse over one block of `block_rows` frames
__global__ void __launch_bounds__(256) k_p2a_partial(const float* __restrict__ p, const P2aJob* __restrict__ jobs,
                                                    int n_jobs, const int32_t* __restrict__ blk_prefix, int bins,
                                                    int block_rows, float realse, double* __restrict__ partial) {
    const int ji = find_utt(blk_prefix, n_jobs, blockIdx.x);
    const P2aJob job = jobs[ji];
    const int64_t r0 = (int64_t)(blockIdx.x - blk_prefix[ji]) * block_rows;
    const int64_t r1 = r0 + block_rows < job.rows ? r0 + block_rows : job.rows;
    const float* __restrict__ src = p + (job.row0 + r0) * bins;
    const int64_t n = (r1 - r0) * bins;
    double s0 = 0.0, s1 = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) {
        const float v = fmaxf(src[i], 0.0f);
        s0 += (double)v;
        s1 += (double)powf(v, realse);
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    __shared__ double red[2][8];
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0;
        for (int w = 0; w < 8; ++w) { t0 += red[0][w]; t1 += red[1][w]; }
        partial[2 * blockIdx.x] = t0;
        partial[2 * blockIdx.x + 1] = t1;
    }
}

// (p_mean / P.mean()) in float32 (:296; the element counts cancel): sequential sum of the block partials.
// `partial` holds blk_prefix[n_jobs] pairs; job j owns pairs [blk_prefix[j], blk_prefix[j+1]).
__global__ void __launch_bounds__(128) k_p2a_scale(int n_jobs, const int32_t* __restrict__ blk_prefix,
                                                  const double* __restrict__ partial, float* __restrict__ scale) {
    const int ji = blockIdx.x * 128 + threadIdx.x;
    if (ji >= n_jobs) return;
    double s0 = 0.0, s1 = 0.0;
    for (int t = blk_prefix[ji]; t < blk_prefix[ji + 1]; ++t) { s0 += partial[2 * t]; s1 += partial[2 * t + 1]; }
    scale[ji] = (float)s0 / (float)s1;
}

// p and amp may be the same buffer (element-wise, in place): no __restrict__ on them
__global__ void __launch_bounds__(256) k_p2a_apply(const float* p, const P2aJob* __restrict__ jobs,
                                                  int n_jobs, const int32_t* __restrict__ prefix, int bins,
                                                  float realse, int use_realse, const float* __restrict__ scale,
                                                  float inv_norm, float* amp) {
    const int ji = find_utt(prefix, n_jobs, blockIdx.x);
    const P2aJob job = jobs[ji];
    const int64_t n = job.rows * bins;
    const int64_t begin = (int64_t)(blockIdx.x - prefix[ji]) * kP2aChunk;
    const int64_t end = begin + kP2aChunk < n ? begin + kP2aChunk : n;
    const float* src = p + job.row0 * bins;
    float* dst = amp + job.row0 * bins;
    const float sc = use_realse ? scale[ji] : 1.0f;
    for (int64_t i = begin + threadIdx.x; i < end; i += 256) {
        float v = fmaxf(src[i], 0.0f);
        if (use_realse) v = sc * powf(v, realse);
        const float db = v * inv_norm - 80.0f;                       // P / P_dB_norm_factor - 80
        dst[i] = exp2f(0.16609640474436813f * db);                   // sqrt(10^(0.1*db)) = 2^(db*log2(10)/20)
    }
}

// ---------------------------------------------------------------------------------------------
// from_power_to_wav epilogue (:301-306): de-emphasis IIR y[n] = x[n] + c*y[n-1] in float64 as a chunked scan,
// then y * (m / mean|y|).
//
// A signal is cut into chunks of kIirChunk samples on its own grid (sample 0 starts a chunk).  Phase 1 computes the
// zero-state response at the end of every chunk (`loc`).  The state entering chunk k is
//     S_k = loc[k-1] + c^256 * loc[k-2] + c^512 * loc[k-3] + ...
// and c^256 is tiny for the reference's coefficients (0.97^256 = 4e-4), so the series is cut after `win` terms
// chosen on the host such that |c|^(256*win) < 1e-40 (12 terms for 0.97): phase 2 evaluates it per chunk with a
// Horner chain over the previous `win` entries of `loc` - no sequential pass over the signal, and a time-chunked run
// only needs the last `win` values of its left neighbour (one small message per boundary).  Every
// quantity is computed on the whole-signal grid, so chunked and unchunked runs are bit-identical.  Coefficients too
// close to 1 for a short window (win > kIirMaxWin) take the sequential carry kernel instead (single GPU only).
// mean|y| uses a canonical order as well: float64 partial per chunk (sequential in the owning thread), sequential sum of
// the `blk_chunks` chunk partials of a block, sequential sum of the blocks.
constexpr int kIirChunk = 256;       // samples per thread-sequential chunk
constexpr int kIirBlock = 128;       // chunks per CTA
constexpr int kIirMaxWin = 64;

struct WavJob {
    int64_t off;      // element offset of the first sample of this job in the wav / out buffers
    int64_t len;      // samples present in this job
    int64_t first;    // whole-signal index of that sample (multiple of kIirChunk); 0 for a whole signal
    int64_t total;    // whole-signal length (renormalisation divides by it)
    int32_t tile0;    // prefix of CTAs (kIirBlock chunks each)
    int32_t chunk0;   // prefix of entries in the `loc` array: this job owns [chunk0, chunk0 + halo + n_chunks)
    int32_t blk0;     // prefix of |y| sum blocks
    int32_t halo;     // entries of `loc` in front of chunk 0 (zeros, or the left neighbour's last chunks)
};

// phase 1: zero-state response at the end of each chunk
template <typename TIN>
__global__ void __launch_bounds__(kIirBlock) k_iir_local(const TIN* __restrict__ x, const WavJob* __restrict__ jobs,
                                                        int n_jobs, const int32_t* __restrict__ prefix, double c,
                                                        double* __restrict__ loc) {
    const int ji = find_utt(prefix, n_jobs, blockIdx.x);
    const WavJob job = jobs[ji];
    const int chunk = (blockIdx.x - prefix[ji]) * kIirBlock + threadIdx.x;
    const int64_t begin = (int64_t)chunk * kIirChunk;
    if (begin >= job.len) return;
    const int64_t end = begin + kIirChunk < job.len ? begin + kIirChunk : job.len;
    const TIN* __restrict__ src = x + job.off;
    double y = 0.0;
    for (int64_t i = begin; i < end; ++i) y = (double)src[i] + c * y;
    loc[job.chunk0 + job.halo + chunk] = y;
}

// fallback phase 2 (|c| close to 1): sequential carry over the chunks of one signal, `loc` becomes the incoming states
__global__ void __launch_bounds__(32) k_iir_carry(const WavJob* __restrict__ jobs, int n_jobs, double c,
                                                 double* __restrict__ loc) {
    const int ji = blockIdx.x * 32 + threadIdx.x;
    if (ji >= n_jobs) return;
    const WavJob job = jobs[ji];
    const int64_t n_chunks = (job.len + kIirChunk - 1) / kIirChunk;
    double cl = 1.0;                                   // c^kIirChunk
    for (int i = 0; i < kIirChunk; ++i) cl *= c;
    double carry = 0.0;                                // state entering chunk k
    double* __restrict__ l = loc + job.chunk0 + job.halo;
    for (int64_t k = 0; k < n_chunks; ++k) {
        const double local = l[k];
        l[k] = carry;                                  // overwrite with the incoming state
        carry = local + cl * carry;                    // every chunk but the last is full
    }
}

// phase 2 + 3: incoming state from the previous `win` chunk responses (win == 0: `loc` already holds the states),
// rerun the chunk from it, write float64, per-chunk sum of |y|
template <typename TIN>
__global__ void __launch_bounds__(kIirBlock) k_iir_apply(const TIN* __restrict__ x, const WavJob* __restrict__ jobs,
                                                        int n_jobs, const int32_t* __restrict__ prefix, double c,
                                                        double cl, int win, const double* __restrict__ loc,
                                                        double* __restrict__ out, double* __restrict__ chunk_abs) {
    const int ji = find_utt(prefix, n_jobs, blockIdx.x);
    const WavJob job = jobs[ji];
    const int chunk = (blockIdx.x - prefix[ji]) * kIirBlock + threadIdx.x;
    const int64_t begin = (int64_t)chunk * kIirChunk;
    if (begin >= job.len) return;
    const int64_t end = begin + kIirChunk < job.len ? begin + kIirChunk : job.len;
    const double* __restrict__ l = loc + job.chunk0 + job.halo + chunk;     // l[-m] = response of chunk - m
    double y;
    if (win == 0) {
        y = l[0];
    } else {
        const int reach = chunk + job.halo < win ? chunk + job.halo : win;  // entries that exist in front of this chunk
        y = 0.0;
        for (int m = reach; m >= 1; --m) y = l[-m] + cl * y;
    }
    const TIN* __restrict__ src = x + job.off;
    double* __restrict__ dst = out + job.off;
    double a = 0.0;
    for (int64_t i = begin; i < end; ++i) {
        y = (double)src[i] + c * y;
        dst[i] = y;
        a += fabs(y);
    }
    chunk_abs[job.chunk0 + job.halo + chunk] = a;
}

// per-block sums of |y|: block b of a job = its chunks [b * blk_chunks, (b+1) * blk_chunks), summed in order
__global__ void __launch_bounds__(128) k_abs_blocks(const WavJob* __restrict__ jobs, int n_jobs,
                                                   const int32_t* __restrict__ blk_prefix, int blk_chunks,
                                                   const double* __restrict__ chunk_abs, double* __restrict__ blk_sum) {
    const int b = blockIdx.x * 128 + threadIdx.x;
    if (b >= blk_prefix[n_jobs]) return;
    const int ji = find_utt(blk_prefix, n_jobs, b);
    const WavJob job = jobs[ji];
    const int64_t n_chunks = (job.len + kIirChunk - 1) / kIirChunk;
    const int64_t k0 = (int64_t)(b - blk_prefix[ji]) * blk_chunks;
    const int64_t k1 = k0 + blk_chunks < n_chunks ? k0 + blk_chunks : n_chunks;
    const double* __restrict__ a = chunk_abs + job.chunk0 + job.halo;
    double s = 0.0;
    for (int64_t k = k0; k < k1; ++k) s += a[k];
    blk_sum[b] = s;
}

// scale[j] = target / (sum of the job's block sums / total samples); blocks summed in order by one thread.
// all_blk != nullptr (time-chunked run, one job): sum the n_all block sums of the WHOLE signal instead.
__global__ void __launch_bounds__(128) k_renorm_scale(const WavJob* __restrict__ jobs, int n_jobs,
                                                     const int32_t* __restrict__ blk_prefix,
                                                     const double* __restrict__ blk_sum, const double* __restrict__ all_blk,
                                                     int n_all, double target, double* __restrict__ scale) {
    const int ji = blockIdx.x * 128 + threadIdx.x;
    if (ji >= n_jobs) return;
    double s = 0.0;
    if (all_blk) for (int t = 0; t < n_all; ++t) s += all_blk[t];
    else for (int t = blk_prefix[ji]; t < blk_prefix[ji + 1]; ++t) s += blk_sum[t];
    scale[ji] = target / (s / (double)jobs[ji].total);
}

// y *= m / mean|y|   (:306)
__global__ void __launch_bounds__(256) k_renorm(const WavJob* __restrict__ jobs, int n_jobs,
                                               const int32_t* __restrict__ prefix, const double* __restrict__ scale,
                                               double* __restrict__ out) {
    const int ji = find_utt(prefix, n_jobs, blockIdx.x);
    const WavJob job = jobs[ji];
    const double sc = scale[ji];
    const int64_t begin = (int64_t)(blockIdx.x - prefix[ji]) * (kIirChunk * kIirBlock);
    const int64_t end = begin + kIirChunk * kIirBlock < job.len ? begin + kIirChunk * kIirBlock : job.len;
    double* __restrict__ dst = out + job.off;
    for (int64_t i = begin + threadIdx.x; i < end; i += 256) dst[i] *= sc;
}

// ---------------------------------------------------------------------------------------------
// standalone calc_preemphasis (:27): y[n] = x[n] - c*x[n-1], float64 out
template <typename TIN>
__global__ void __launch_bounds__(256) k_preemph(const TIN* __restrict__ x, int64_t n, double c,
                                                double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const double prev = i > 0 ? (double)x[i - 1] : 0.0;
    out[i] = (double)x[i] - c * prev;
}

// [rows][cols] -> [cols][rows] float32 (float64 or float32 source)
template <typename T>
__global__ void __launch_bounds__(256) k_transpose(const T* __restrict__ src, int64_t rows, int64_t cols,
                                                  float* __restrict__ dst) {
    __shared__ float tile[32][33];
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) {
        const int64_t r = r0 + k, c = c0 + tx;
        tile[k][tx] = (r < rows && c < cols) ? (float)src[r * cols + c] : 0.f;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int64_t c = c0 + k, r = r0 + tx;
        if (c < cols && r < rows) dst[c * rows + r] = tile[tx][k];
    }
}

}  // namespace scdsp
