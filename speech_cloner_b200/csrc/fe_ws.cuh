// Warp-specialised persistent pass A of the front-end (fast path: n_fft = 400, hop = 80).
//
// Same arithmetic as k_fe_pass_a (audio_lib.py:125-172: gain -> pre-emphasis -> reflect pad -> window
// -> rFFT-400 -> |X|^2 -> raw power (dB taken in pass B), sparse Slaney mel -> raw mel dB, utterance max / min), but the
// three kinds of work run on different warps of one CTA per SM and overlap tile against tile:
//
//   warp 15      PREP      stages the raw samples of tile i+2 with one TMA bulk copy (cp.async.bulk ->
//                          mbarrier complete_tx; reflect-padded edge tiles are gathered with cp.async
//                          arriving on the same mbarrier), and turns tile i into gain-scaled,
//                          pre-emphasised float64 samples (each sample once, not once per frame)
//   warps 0-7    FFT       two independent groups of 6 units x 20 threads (12 frames each): windowed
//                          real FFT-400 in float64, |X|^2 into a double-buffered power tile.  Only
//                          FP64-pipe and shared-memory work: no global access, no MUFU
//   warps 8-14   EPILOGUE  |X|^2 out with coalesced 128-bit stores (pass B takes the logarithm), the sparse mel
//                          filterbank with lane = frame (two bands per step, 4 taps per block, weights as broadcast
//                          128-bit loads), raw mel dB, utterance max / min.  FP32 / MUFU / LSU work that runs beside
//                          the FFT warps
//
// Hand-offs are mbarriers (full / empty pairs, parity from the tile counter), the two FFT groups
// use one named barrier each, so no warp ever waits for a role it does not depend on.
#pragma once
#include "fe_kernels.cuh"

namespace scdsp {

constexpr int kWsUnits = 12;                                 // frame pairs per tile
constexpr int kWsFrames = 2 * kWsUnits;                      // 24
constexpr int kWsSpan = kHop * (kWsFrames - 1) + kNfft;      // 2240 samples a tile touches
constexpr int kWsRaw = kWsSpan + 8;                          // raw[i + 4] = y[reflect(q0 + i)], 16-byte multiple
constexpr int kWsGroupUnits = 6;
constexpr int kWsGroupThreads = 128;                         // 120 working + 8 idle
constexpr int kWsFftThreads = 2 * kWsGroupThreads;
constexpr int kWsEpiWarps = 7;                               // 8 FFT + 7 epilogue + 1 prep = 16 warps: 128 registers per thread
constexpr int kWsEpiThreads = 32 * kWsEpiWarps;
constexpr int kWsThreads = kWsFftThreads + kWsEpiThreads + 32;   // 512
constexpr int kWsMaxTaps = 2048;                             // padded filterbank taps (floats) the kernel stages in shared memory
// Optional register reallocation between the roles (setmaxnreg, per 4-warp group): the kernel launches with 128
// registers per thread (16 warps = the whole register file); the two epilogue / prep warp groups give registers back
// and the two FFT warp groups take them: 8*32*168 + 8*32*88 = 65536.  Measured (B200): no gain for the FFT role (it is
// not register-ILP bound) and a slower epilogue (0.2096 -> 0.2124 ms), so it is off by default.
#ifndef SC_WS_SETMAXNREG
#define SC_WS_SETMAXNREG 0
#endif
#ifndef SC_WS_REG_FFT
#define SC_WS_REG_FFT 168
#endif
#ifndef SC_WS_REG_EPI
#define SC_WS_REG_EPI 88
#endif
constexpr int kWsRawSlots = 3;
constexpr int kWsUnitSlots = 426;                            // complex slots per unit: 20 rows x 21 (+6): unit stride = 2 (mod 8) in 16-byte words
constexpr int kWsDescRing = 8;

// one record per tile, written by k_fe_setup
struct WsTile {
    int64_t sample_off;     // first sample of the utterance in the packed buffer
    int64_t L;              // utterance length
    int64_t frame_row;      // first output row of the tile
    int32_t u, t0, nfr, edge;
};

// Mel filterbank of the epilogue (built by the host): every epilogue warp owns a contiguous band range and
// walks it two bands at a time.  Pair record: x / y = first tap (bin) of band A / B, z = float4 blocks (both
// bands zero padded to the same count), w = offset of the pair's weights in float4 units, laid out
// A0 B0 A1 B1 ...; band = rising slope over interval b + falling slope over interval b+1 (audio_lib.py:160-169).
constexpr int kWsMaxPairs = kMaxMels / 2 + kWsEpiWarps;
struct WsMelParam {
    int32_t chunk[kWsEpiWarps + 1];       // band range per epilogue warp, cost balanced
    int32_t pair0[kWsEpiWarps + 1];       // first pair record of each warp
    int32_t n_mels;
    int32_t n_taps;                       // floats in the weight table (multiple of 8, <= kWsMaxTaps)
};

// One launch that writes every per-call table of the fast front-end path: the sub-tree records of the |y| sum,
// the tiles of k_fe_pass_a_ws and the tiles of k_fe_pass_b3 (the per-utterance statistics are reset by
// k_gain_finalize).  Thread id space: [0, n_abs) | [n_abs, n_abs + n_ws) | [.., + n_b3).
struct SetupArgs {
    const int32_t* pre_abs; const int32_t* pre_ws; const int32_t* pre_b3;
    const int64_t* heap_off;
    int32_t n_abs, n_ws, n_b3;
    int32_t ws_frames;            // frames per pass A tile: kWsFrames (k_fe_pass_a_ws) or kSpFrames (k_fe_fft)
    AbsRec* abs_out; WsTile* ws_out; B3Tile* b3_out;
};
__global__ void k_fe_setup(Ragged rg, SetupArgs sa) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < sa.n_abs) {
        const int u = find_utt(sa.pre_abs, rg.n_utts, i);
        const int idx = i - sa.pre_abs[u];
        const int64_t len = rg.sample_len[u];
        const int D = abs_depth(len);
        int64_t start = 0, n = len;
        for (int b = D - 1; b >= 0; --b) {
            int64_t n2 = n / 2;
            n2 -= n2 % 8;
            if ((idx >> b) & 1) { start += n2; n -= n2; } else { n = n2; }
        }
        AbsRec r;
        r.src_off = rg.sample_off[u] + start;
        r.heap_pos = sa.heap_off[u] + (int64_t(1) << D) + idx;
        r.n = (int)n; r.pad = 0;
        sa.abs_out[i] = r;
        return;
    }
    i -= sa.n_abs;
    if (i < sa.n_ws) {
        const int u = find_utt(sa.pre_ws, rg.n_utts, i);
        WsTile t;
        t.u = u;
        t.t0 = (i - sa.pre_ws[u]) * sa.ws_frames;
        t.nfr = min(sa.ws_frames, rg.frame_cnt[u] - t.t0);
        t.L = rg.sample_len[u];
        t.sample_off = rg.sample_off[u];
        t.frame_row = rg.frame_off[u] + t.t0;
        const int64_t q0 = (int64_t)t.t0 * kHop - kNfft / 2;
        t.edge = !(q0 - 4 >= 0 && q0 + (kHop * (sa.ws_frames - 1) + kNfft) + 4 <= t.L);
        sa.ws_out[i] = t;
        return;
    }
    i -= sa.n_ws;
    if (i < sa.n_b3) {
        const int u = find_utt(sa.pre_b3, rg.n_utts, i);
        B3Tile t;
        t.u = u;
        t.t0 = (i - sa.pre_b3[u]) * kB3Frames;
        t.T = rg.frame_cnt[u];
        t.frame_off = rg.frame_off[u];
        t.pad = 0;
        sa.b3_out[i] = t;
    }
}

// ---- mbarrier / bulk-copy primitives (PTX)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cp_async(uint64_t* b) {   // arrives once this thread's earlier cp.async have landed
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
    return ok != 0;
}
// the same probe, but the hardware may park the warp for up to `ns` nanoseconds waiting for the phase to complete
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* b, unsigned parity, unsigned ns) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
// Waiting warps are parked by the hardware (suspend-time hint) so that they do not take issue slots from the
// warps they wait for.  A lost arrival would hang the GPU: trap instead after ~1 s so the host sees an error.
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
    if (mbar_try_wait(b, parity)) return;
    unsigned spins = 0;
    while (!mbar_try_wait_hint(b, parity, 4000)) {
        if (++spins > (1u << 20)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Optional role timing (build with -DSC_WS_DEBUG): block 0 prints, per role, the cycles spent inside each wait.
#ifdef SC_WS_DEBUG
#define WS_TIMED(acc, stmt) do { const long long t_ = clock64(); stmt; acc += clock64() - t_; } while (0)
#else
#define WS_TIMED(acc, stmt) do { stmt; } while (0)
#endif

template <typename R>
struct WsSmem {
    cx<R> slots[kWsUnits * kWsUnitSlots];             // step-1 -> step-2 exchange, one region per unit
    R span[2][kWsSpan];                               // gain-scaled, pre-emphasised samples of a tile
    R win[kNfft];
    cx<R> w400[sizeof(R) == 8 ? kNfft : 1];
    alignas(16) float raw[kWsRawSlots][kWsRaw];       // TMA / cp.async landing ring
    alignas(16) float power[2][kWsFrames * kBins + 64];  // |X|^2, row = frame (+ pad: zero-weight taps may read past the last row)
    alignas(16) float wt[kWsMaxTaps];                 // padded mel weights
    int4 prec[kWsMaxPairs + 1];                       // band-pair records (+1: prefetch past the end)
    WsTile desc[kWsDescRing];
    alignas(8) uint64_t bar_raw_full[kWsRawSlots];
    uint64_t bar_span_full[2], bar_span_empty[2], bar_pw_full[2], bar_pw_empty[2];
    // followed by mel_s[2][kWsFrames][n_mels | 1]
};

// NB float4 blocks of two bands at once: all loads first, then two independent FMA chains
template <int NB>
__device__ __forceinline__ void ws_mel_pair(const float* __restrict__ pa, const float* __restrict__ pb,
                                            const float4* __restrict__ ww, float& acc_a, float& acc_b) {
    float4 wa[NB], wb[NB];
    float a[4 * NB], b[4 * NB];
#pragma unroll
    for (int q = 0; q < NB; ++q) { wa[q] = ww[2 * q]; wb[q] = ww[2 * q + 1]; }
#pragma unroll
    for (int q = 0; q < 4 * NB; ++q) { a[q] = pa[q]; b[q] = pb[q]; }
#pragma unroll
    for (int q = 0; q < NB; ++q) {
        acc_a = fmaf(wa[q].x, a[4 * q], acc_a);     acc_b = fmaf(wb[q].x, b[4 * q], acc_b);
        acc_a = fmaf(wa[q].y, a[4 * q + 1], acc_a); acc_b = fmaf(wb[q].y, b[4 * q + 1], acc_b);
        acc_a = fmaf(wa[q].z, a[4 * q + 2], acc_a); acc_b = fmaf(wb[q].z, b[4 * q + 2], acc_b);
        acc_a = fmaf(wa[q].w, a[4 * q + 3], acc_a); acc_b = fmaf(wb[q].w, b[4 * q + 3], acc_b);
    }
}

template <typename R>
__global__ void __launch_bounds__(kWsThreads, 1)
k_fe_pass_a_ws(const float* __restrict__ wav, const WsTile* __restrict__ tiles, int total_tiles, FeTables tb, FeParams prm,
               UttStat* __restrict__ stat, float* __restrict__ pdb_out, float* __restrict__ mel_raw,
               const int4* __restrict__ mel_brec, const float* __restrict__ mel_wt, WsMelParam mp) {
    using SM = WsSmem<R>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    float* mel_s = reinterpret_cast<float*>(smem_raw + sizeof(SM));
    constexpr bool kF64 = sizeof(R) == 8;
    const int tid = threadIdx.x;
    const int G = gridDim.x;
    const int n_mels = mp.n_mels;
    const int mel_ld = n_mels | 1;

    // ---- one-time setup: tables, barriers
    if (kF64) {
        for (int i = tid; i < kNfft; i += kWsThreads) {
            sm.win[i] = (R)(2.0 * tb.win_half_d[i]);
            sm.w400[i] = mk<R>((R)tb.w400_d[i].x, (R)tb.w400_d[i].y);
        }
    } else {
        for (int i = tid; i < kNfft; i += kWsThreads) sm.win[i] = (R)(2.0f * tb.win_half[i]);
    }
    for (int i = tid; i < mp.n_taps; i += kWsThreads) sm.wt[i] = mel_wt[i];
    for (int i = tid; i <= kWsMaxPairs; i += kWsThreads) sm.prec[i] = mel_brec[i];
    if (tid < 128) sm.power[tid >> 6][kWsFrames * kBins + (tid & 63)] = 0.f;
    if (tid == 0) {
        for (int s = 0; s < kWsRawSlots; ++s) mbar_init(&sm.bar_raw_full[s], 32);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sm.bar_span_full[b], 32);
            mbar_init(&sm.bar_span_empty[b], kWsFftThreads);
            mbar_init(&sm.bar_pw_full[b], kWsFftThreads);
            mbar_init(&sm.bar_pw_empty[b], kWsEpiThreads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    __syncthreads();

    if (tid >= kWsFftThreads + kWsEpiThreads) {
        // =========================================== PREP warp ===========================================
#if SC_WS_SETMAXNREG
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SC_WS_REG_EPI));
#endif
        const int lane = tid & 31;
        const double c = prm.pre_emphasis;
        // stage(): publish the descriptor of local tile n and start the asynchronous copy of its raw samples into
        // ring slot n % 3; returns the utterance gain (loaded here, first used two tiles later)
        auto stage = [&](int n, const WsTile& t) -> float {
            if (blockIdx.x + n * G >= total_tiles) return 0.f;
            const int s = n % kWsRawSlots;
            if (lane == 0) sm.desc[n % kWsDescRing] = t;
            const float gain = __ldg(&stat[t.u].gain);
            const float* __restrict__ y = wav + t.sample_off;
            const int64_t q0 = (int64_t)t.t0 * kHop - kNfft / 2;
            float* dst = sm.raw[s];
            const float* __restrict__ src = y + (q0 - 4);
            if (!t.edge && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
                fence_proxy_async();                  // the slot was last read through the generic proxy
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_expect_tx(&sm.bar_raw_full[s], kWsRaw * 4);
                    bulk_g2s(dst, src, kWsRaw * 4, &sm.bar_raw_full[s]);
                } else {
                    mbar_arrive(&sm.bar_raw_full[s]);
                }
            } else {
                if (!t.edge) {
                    for (int i = lane; i < kWsRaw; i += 32) cp_async4(dst + i, src + i);
                } else {
                    // np.pad(.., 'reflect') (audio_lib.py:147) as a gather; the host guarantees one reflection per side
                    for (int i = lane; i < kWsRaw; i += 32) {
                        const int64_t q = q0 - 4 + i;
                        int64_t r = q < 0 ? -q : (q > t.L - 1 ? 2 * (t.L - 1) - q : q);
                        r = r < 0 ? 0 : (r > t.L - 1 ? t.L - 1 : r);      // only reachable by samples of frames past T
                        cp_async4(dst + i, y + r);
                    }
                }
                mbar_arrive_cp_async(&sm.bar_raw_full[s]);
            }
            __syncwarp();
            return gain;
        };
        // Tiles are walked from the LAST utterance to the first: the |y| pass before this kernel read the batch front to
        // back, so its tail is what L2 still holds, and pass B (front to back) then starts on the raw tiles this kernel
        // wrote last - they are re-read from L2 and overwritten in place before they are ever written back to DRAM.
        auto fetch = [&](int n) -> WsTile {
            const int tile = blockIdx.x + n * G;
            WsTile t;
            t.sample_off = 0; t.L = 1; t.frame_row = 0; t.u = 0; t.t0 = 0; t.nfr = 0; t.edge = 0;
            if (tile < total_tiles) t = tiles[total_tiles - 1 - tile];
            return t;
        };
        long long w0 = 0, w1 = 0, w2 = 0, w3 = 0;
        const long long t_begin = clock64();
        float g0 = stage(0, fetch(0));
        float g1 = stage(1, fetch(1));
        WsTile rec = fetch(2);
        for (int i = 0; blockIdx.x + i * G < total_tiles; ++i) {
            const WsTile rec_next = fetch(i + 3);          // consumed by the next iteration's stage()
            const float g2 = stage(i + 2, rec);
            const int s = i % kWsRawSlots, b = i & 1, k = i >> 1;
            WS_TIMED(w0, mbar_wait(&sm.bar_raw_full[s], (i / kWsRawSlots) & 1));
            const WsTile t = sm.desc[i % kWsDescRing];
            const float gain = g0;
            WS_TIMED(w1, mbar_wait(&sm.bar_span_empty[b], (k & 1) ^ 1));
            const float* __restrict__ raw = sm.raw[s] + 4;
            R* __restrict__ sp = sm.span[b];
            if (!t.edge) {
                // gain in float32 (:126), pre-emphasis in float64 (:27); two samples per lane and step
                // (loads of a batch first: the compiler keeps shared-memory loads behind earlier shared-memory stores)
                constexpr int kBatch = 7;
                static_assert(kWsSpan % (64 * kBatch) == 0, "span must split into whole batches");
                for (int e0 = 2 * lane; e0 < kWsSpan; e0 += 64 * kBatch) {
                    float pm[kBatch];
                    float2 cu[kBatch];
#pragma unroll
                    for (int q = 0; q < kBatch; ++q) {
                        pm[q] = raw[e0 + 64 * q - 1];
                        cu[q] = *reinterpret_cast<const float2*>(raw + e0 + 64 * q);
                    }
#pragma unroll
                    for (int q = 0; q < kBatch; ++q) {
                        const float p0 = gain * pm[q], c0 = gain * cu[q].x, c1 = gain * cu[q].y;
                        cx<R> v;
                        v.x = (R)((double)c0 - c * (double)p0);
                        v.y = (R)((double)c1 - c * (double)c0);
                        *reinterpret_cast<cx<R>*>(sp + e0 + 64 * q) = v;
                    }
                }
            } else {
                // y[r - 1] of a reflected sample is its right-hand neighbour in the padded order; y[-1] = 0 (:27)
                const int64_t q0 = (int64_t)t.t0 * kHop - kNfft / 2;
                constexpr int kBatchE = 7;
                static_assert(kWsSpan % (32 * kBatchE) == 0, "span must split into whole batches");
                for (int e0 = lane; e0 < kWsSpan; e0 += 32 * kBatchE) {
                    float lo[kBatchE], mid[kBatchE], hi[kBatchE];
#pragma unroll
                    for (int q = 0; q < kBatchE; ++q) {
                        lo[q] = raw[e0 + 32 * q - 1];
                        mid[q] = raw[e0 + 32 * q];
                        hi[q] = raw[e0 + 32 * q + 1];
                    }
#pragma unroll
                    for (int q = 0; q < kBatchE; ++q) {
                        const int64_t pos = q0 + e0 + 32 * q;
                        const float cur = gain * mid[q];
                        const float nb = (pos < 0 || pos > t.L - 1) ? hi[q] : lo[q];
                        const float prev = pos == 0 ? 0.0f : gain * nb;
                        sp[e0 + 32 * q] = (R)((double)cur - c * (double)prev);
                    }
                }
            }
            mbar_arrive(&sm.bar_span_full[b]);
            g0 = g1; g1 = g2; rec = rec_next;
        }
#ifdef SC_WS_DEBUG
        if (blockIdx.x == 0 && lane == 0) printf("prep: total %lld raw_full %lld span_empty %lld\n", clock64() - t_begin, w0, w1);
#endif
        (void)w2; (void)w3; (void)t_begin;
        return;
    }

    if (tid < kWsFftThreads) {
        // =========================================== FFT warps ===========================================
#if SC_WS_SETMAXNREG
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SC_WS_REG_FFT));
#endif
        const int g = tid >> 7;                           // group 0 / 1: frames [12 g, 12 g + 12) of the tile
        const int lt = tid & 127;
        const int lane = tid & 31;
        // step-1 mapping (unit, j): threads 0..95 are 6 units x (j = 0..15), so a half-warp reads 16 consecutive
        // samples / window values / slots of ONE unit; threads 96..127 hold j = 16..19 in groups of 4, paired so
        // that the two units of a quarter-warp are two apart (their slot regions differ by 4 sixteen-byte words)
        int ul1, j;
        bool s1_on = true;
        if (lt < 96) {
            ul1 = lt >> 4;
            j = lt & 15;
        } else {
            const int q = (lt - 96) >> 2;                 // 0..7 -> unit 0 2 1 3 4 - 5 -
            ul1 = q == 0 ? 0 : q == 1 ? 2 : q == 2 ? 1 : q == 3 ? 3 : q == 4 ? 4 : 5;
            s1_on = q != 5 && q != 7;
            j = 16 + (lt & 3);
        }
        const int unit1 = g * kWsGroupUnits + ul1;
        // step-2 task.  Logical warps 0, 1 run 32 one-frame columns each; logical warps 2, 3 run 22 one-frame
        // columns + the 6 packed columns c = 0 (warp 2) or c = 10 (warp 3).  Group 1 swaps the pairs so that each
        // SM sub-partition hosts one light and one heavy warp.  One-frame column task gt = 18 * unit + row reads
        // slot row `row`; with the unit stride of 426 its 16-byte word address is 5 * gt (mod 8): any 8 consecutive
        // tasks hit 8 different bank groups.
        const int lwl = ((lt >> 5) ^ (g << 1)) & 3;
        int ul2 = 0, c2 = -1, row2 = 0;
        {
            int gt = -1;
            if (lwl < 2) gt = 32 * lwl + lane;
            else if (lane < 22) gt = 64 + 22 * (lwl - 2) + lane;
            else if (lane < 22 + kWsGroupUnits) { ul2 = lane - 22; c2 = lwl == 2 ? 0 : 10; row2 = lwl == 2 ? 18 : 19; }
            if (gt >= 0) {
                ul2 = gt / 18;
                row2 = gt - ul2 * 18;
                c2 = row2 < 9 ? row2 + 1 : row2 + 2;
            }
        }
        const int unit2 = g * kWsGroupUnits + ul2;
        // twiddles: W400^j and W400^(10 j) stay in registers, W400^(j k1) is built by repeated multiplication
        // (rounding of the chain is ~1e-15 in float64, ~5e-7 in float32: the float32 mode reads its table instead)
        cx<R> w1, w10;
        typename FeTw<R>::type tw;
        if (kF64) {
            w1 = sm.w400[j];
            w10 = sm.w400[10 * j];
        } else {
            tw.load(reinterpret_cast<const cx<R>*>(tb.w400), j);
        }
        cx<R>* slot_col = sm.slots + unit1 * kWsUnitSlots + j;
        const cx<R>* slot_row = sm.slots + unit2 * kWsUnitSlots + row2 * kSlotLd;
        const int bar_id = 1 + g;
        long long w0 = 0, w1t = 0, w2 = 0, w3 = 0;
        const long long t_begin = clock64();
        for (int i = 0; blockIdx.x + i * G < total_tiles; ++i) {
            const int b = i & 1, k = i >> 1;
            WS_TIMED(w0, mbar_wait(&sm.bar_span_full[b], k & 1));
            if (s1_on) {
                const R* __restrict__ src = sm.span[b] + unit1 * (2 * kHop) + j;
                R s[24];
#pragma unroll
                for (int m = 0; m < 24; ++m) s[m] = src[20 * m];
                R xa[20], xb[20];
#pragma unroll
                for (int n1 = 0; n1 < 20; ++n1) {
                    const R w = sm.win[20 * n1 + j];
                    xa[n1] = s[n1] * w;
                    xb[n1] = s[n1 + 4] * w;
                }
                // two real 20-point DFTs, twiddle, store: column c goes to slot row (c - 1) for c = 1..9,
                // (c - 2) for c = 11..19, the packed columns c = 0 / 10 to rows 18 / 19
                cx<R> ya[11], yb[11];
                rdft20_fwd(xa, ya);
                rdft20_fwd(xb, yb);
                slot_col[18 * kSlotLd] = mk<R>(ya[0].x, yb[0].x);
                if (kF64) {
                    slot_col[19 * kSlotLd] = cmul(mk<R>(ya[10].x, yb[10].x), w10);
                    cx<R> w = w1;
#pragma unroll
                    for (int k1 = 1; k1 < 10; ++k1) {
                        slot_col[(k1 - 1) * kSlotLd] = cmul(ya[k1], w);
                        slot_col[(k1 + 8) * kSlotLd] = cmul(yb[k1], w);
                        if (k1 < 9) w = cmul(w, w1);
                    }
                } else {
                    slot_col[19 * kSlotLd] = cmul(mk<R>(ya[10].x, yb[10].x), tw.get10());
#pragma unroll
                    for (int k1 = 1; k1 < 10; ++k1) {
                        const cx<R> w = tw.get(k1);
                        slot_col[(k1 - 1) * kSlotLd] = cmul(ya[k1], w);
                        slot_col[(k1 + 8) * kSlotLd] = cmul(yb[k1], w);
                    }
                }
            }
            WS_TIMED(w1t, asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kWsGroupThreads) : "memory"));
            mbar_arrive(&sm.bar_span_empty[b]);
            WS_TIMED(w2, mbar_wait(&sm.bar_pw_empty[b], (k & 1) ^ 1));
            if (c2 >= 0) {
                cx<R> v[20];
                fwd_step2(v, slot_row);
                float* pa = sm.power[b] + (2 * unit2) * kBins;
                store_power(v, c2, pa, pa + kBins);
            }
            mbar_arrive(&sm.bar_pw_full[b]);
            WS_TIMED(w3, asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kWsGroupThreads) : "memory"));
        }
#ifdef SC_WS_DEBUG
        if (blockIdx.x == 0 && (tid & 31) == 0)
            printf("fft warp %d: total %lld span_full %lld bar1 %lld pw_empty %lld bar2 %lld\n", tid >> 5, clock64() - t_begin, w0, w1t, w2, w3);
#endif
        (void)w0; (void)w1t; (void)w2; (void)w3; (void)t_begin;
        return;
    }

    // ============================================ EPILOGUE warps ============================================
#if SC_WS_SETMAXNREG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SC_WS_REG_EPI));
#endif
    // Few warps do this work, so every phase is written for instruction-level parallelism: fixed trip counts
    // with predication (loads issue back to back), and a branch-light mel walk driven by per-bin records.
    {
        const int et = tid - kWsFftThreads;
        const int lane = et & 31, ew = et >> 5;
        const int mb = mp.chunk[ew], me = mp.chunk[ew + 1];
        const int pr0 = mp.pair0[ew], pr1 = mp.pair0[ew + 1];
        const float kInf = __int_as_float(0x7f800000);
        long long w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0;
        const long long t_begin = clock64();
        for (int i = 0; blockIdx.x + i * G < total_tiles; ++i) {
            const int b = i & 1, k = i >> 1;
            WS_TIMED(w0, mbar_wait(&sm.bar_pw_full[b], k & 1));
#ifdef SC_WS_DEBUG
            const long long t_a = clock64();
#endif
#ifdef SC_WS_NO_EPI      // experiment: how fast is the prep + FFT pipeline alone?
            mbar_arrive(&sm.bar_pw_empty[b]);
            continue;
#endif
            const WsTile t = sm.desc[i % kWsDescRing];
            const float* __restrict__ P = sm.power[b];
            float* __restrict__ ms = mel_s + b * (kWsFrames * mel_ld);
            const int nfr = t.nfr;
            float p_max = 0.f, p_min = kInf, m_max = 0.f, m_min = kInf;
            // ---- raw power dB (:157 without the top_db clip), coalesced
            {
                float* __restrict__ dst = pdb_out + t.frame_row * kBins;
                const int n = nfr * kBins;
                int head = 0;
                if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(P);
                    float4* __restrict__ d4 = reinterpret_cast<float4*>(dst);
                    const int n4 = n >> 2;
                    constexpr int kIt = (kWsFrames * kBins / 4 + kWsEpiThreads - 1) / kWsEpiThreads;   // 6
                    float4 p[kIt];
#pragma unroll
                    for (int it = 0; it < kIt; ++it) {
                        const int e = et + kWsEpiThreads * it;
                        p[it] = e < n4 ? p4[e] : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int it = 0; it < kIt; ++it) {
                        const int e = et + kWsEpiThreads * it;
                        if (e < n4) {
                            const float4 q = p[it];
                            p_max = fmaxf(fmaxf(p_max, fmaxf(q.x, q.y)), fmaxf(q.z, q.w));
                            p_min = fminf(fminf(p_min, fminf(q.x, q.y)), fminf(q.z, q.w));
                            d4[e] = make_float4(pdb_store(q.x), pdb_store(q.y), pdb_store(q.z), pdb_store(q.w));
                        }
                    }
                    head = n4 << 2;
                }
                for (int e = head + et; e < n; e += kWsEpiThreads) {
                    const float p = P[e];
                    p_max = fmaxf(p_max, p);
                    p_min = fminf(p_min, p);
                    dst[e] = pdb_store(p);
                }
            }
#ifdef SC_WS_DEBUG
            const long long t_b = clock64(); w1 += t_b - t_a;
#endif
            // ---- sparse Slaney mel (:160-169): lane = frame, two bands per step; mel POWER goes to mel_s
            if (lane < kWsFrames) {
                const float* __restrict__ prow = P + lane * kBins;
                float* __restrict__ mrow = ms + lane * mel_ld + mb;
                int4 d = sm.prec[pr0];
                for (int pi = pr0; pi < pr1; ++pi) {
                    const int4 nxt = sm.prec[pi + 1];          // next record while this pair computes
                    const float* __restrict__ pa = prow + d.x;
                    const float* __restrict__ pb = prow + d.y;
                    const float4* __restrict__ ww = reinterpret_cast<const float4*>(sm.wt) + d.w;
                    float acc_a = 0.f, acc_b = 0.f;
                    int nb = d.z;
                    while (nb > 4) {                           // bands wider than 16 bins (not at the hp settings)
                        ws_mel_pair<4>(pa, pb, ww, acc_a, acc_b);
                        pa += 16; pb += 16; ww += 8; nb -= 4;
                    }
                    switch (nb) {
                        case 4: ws_mel_pair<4>(pa, pb, ww, acc_a, acc_b); break;
                        case 3: ws_mel_pair<3>(pa, pb, ww, acc_a, acc_b); break;
                        case 2: ws_mel_pair<2>(pa, pb, ww, acc_a, acc_b); break;
                        case 1: ws_mel_pair<1>(pa, pb, ww, acc_a, acc_b); break;
                        default: break;
                    }
                    const int o = 2 * (pi - pr0);
                    mrow[o] = acc_a;
                    if (mb + o + 1 < me) mrow[o + 1] = acc_b;
                    d = nxt;
                }
            }
#ifdef SC_WS_DEBUG
            w2 += clock64() - t_b;
#endif
            mbar_arrive(&sm.bar_pw_empty[b]);                  // the power tile is free again
            WS_TIMED(w3, asm volatile("bar.sync 3, %0;" ::"n"(kWsEpiThreads) : "memory"));
#ifdef SC_WS_DEBUG
            const long long t_c = clock64();
#endif
            // ---- raw amplitude_to_db (:172) of the mel power, coalesced rows (at most 4 rows per warp: all loads first)
            {
                constexpr int kRows = (kWsFrames + kWsEpiWarps - 1) / kWsEpiWarps;
                constexpr int kCols = kMaxMels / 32;
                float* __restrict__ dst = mel_raw + t.frame_row * n_mels + lane;
                const float* __restrict__ src = ms + lane;
                float v[kRows][kCols];
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    const int f = ew + kWsEpiWarps * r;
#pragma unroll
                    for (int q = 0; q < kCols; ++q)
                        v[r][q] = (f < nfr && lane + 32 * q < n_mels) ? src[f * mel_ld + 32 * q] : 1.0f;
                }
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    const int f = ew + kWsEpiWarps * r;
#pragma unroll
                    for (int q = 0; q < kCols; ++q) {
                        if (f < nfr && lane + 32 * q < n_mels) {
                            m_max = fmaxf(m_max, v[r][q]);
                            m_min = fminf(m_min, v[r][q]);
                            dst[f * n_mels + 32 * q] = 2.0f * db10(fmaxf(v[r][q], 1e-5f));
                        }
                    }
                }
            }
#ifdef SC_WS_DEBUG
            w4 += clock64() - t_c;
#endif
            // ---- utterance max / min: values are >= +0, so their bit patterns order like unsigned integers
            {
                const unsigned u_pmax = __reduce_max_sync(0xffffffffu, __float_as_uint(p_max));
                const unsigned u_pmin = __reduce_min_sync(0xffffffffu, __float_as_uint(p_min));
                const unsigned u_mmax = __reduce_max_sync(0xffffffffu, __float_as_uint(m_max));
                const unsigned u_mmin = __reduce_min_sync(0xffffffffu, __float_as_uint(m_min));
                if (lane == 0) {
                    UttStat* su = stat + t.u;
                    atomicMax(&su->p_max, u_pmax);
                    atomicMin(&su->p_min, u_pmin);
                    atomicMax(&su->m_max, u_mmax);
                    atomicMin(&su->m_min, u_mmin);
                }
            }
        }
#ifdef SC_WS_DEBUG
        if (blockIdx.x == 0 && lane == 0)
            printf("epi warp %d: total %lld pw_full %lld pdb %lld mel %lld bar %lld copy %lld\n", ew, clock64() - t_begin, w0, w1, w2, w3, w4);
#endif
        (void)w0; (void)w1; (void)w2; (void)w3; (void)w4; (void)t_begin;
    }
}

}  // namespace scdsp
