// Pass A split in two kernels (fast path: n_fft = 400, hop = 80), opt-in alternative to k_fe_pass_a_ws.
//
// k_fe_pass_a_ws is bounded by its FFT role: two groups of FFT warps per SM reach 45 % of the FP64 pipe and shared
// memory (213 KB) has no room for a third group next to the epilogue's power / mel tiles.  Here the epilogue leaves:
//
//   k_fe_fft<R>   persistent, one CTA per SM = THREE independent pipelines of (1 prep warp + 4 FFT warps = 6 units = 12
//                 frames per tile).  Same prep / step 1 / step 2 arithmetic as k_fe_pass_a_ws; |X|^2 goes through the
//                 pipeline's own (then idle) slot region to the power-dB buffer with coalesced 128-bit stores, and the
//                 utterance max / min of the power are reduced on the way.
//   k_fe_mel      streaming kernel over the stored |X|^2: 32-frame tiles, lane = frame, the pair-record Slaney
//                 filterbank of the warp-specialised kernel, raw mel dB + utterance max / min of the mel power.
//
// It reads the power back once (165 MB more DRAM / L2 traffic per 205 k frames) in exchange for 1.5x the FFT
// thread-level parallelism.  audio_lib.py:125-172, same numbers as the fused kernel bit for bit.
#pragma once
#include "fe_ws.cuh"

namespace scdsp {

constexpr int kSpPipes = 3;
constexpr int kSpUnits = 6;                                   // frame pairs per pipeline tile
constexpr int kSpFrames = 2 * kSpUnits;                       // 12
constexpr int kSpSpan = kHop * (kSpFrames - 1) + kNfft;       // 1280
constexpr int kSpRaw = kSpSpan + 8;                           // 1288 floats: 16-byte multiple
constexpr int kSpRawSlots = 2;
constexpr int kSpFftThreads = 128 * kSpPipes;                 // warps 0..11
constexpr int kSpThreads = kSpFftThreads + 128;               // + a warp group holding one prep warp per pipeline (and an idle warp) = 512
constexpr int kSpRegFft = 152, kSpRegPrep = 56;               // setmaxnreg: 12*32*152 + 4*32*56 = 65536
constexpr int kSpDescRing = 8;                               // prep may be 4 tiles ahead of a group that still reads its descriptor

template <typename R>
struct SpPipe {
    cx<R> slots[kSpUnits * kWsUnitSlots];             // step-1 -> step-2 exchange; afterwards the |X|^2 staging tile
    R span[2][kSpSpan];
    alignas(16) float raw[kSpRawSlots][kSpRaw];
    WsTile desc[kSpDescRing];
    alignas(8) uint64_t bar_raw_full[kSpRawSlots];
    uint64_t bar_span_full[2], bar_span_empty[2];
};
template <typename R>
struct SpSmem {
    SpPipe<R> pipe[kSpPipes];
    R win[kNfft];
    cx<R> w400[sizeof(R) == 8 ? kNfft : 1];
};

template <typename R>
__global__ void __launch_bounds__(kSpThreads, 1)
k_fe_fft(const float* __restrict__ wav, const WsTile* __restrict__ tiles, int total_tiles, FeTables tb, FeParams prm,
         UttStat* __restrict__ stat, float* __restrict__ pdb_out) {
    using SM = SpSmem<R>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    constexpr bool kF64 = sizeof(R) == 8;
    const int tid = threadIdx.x;
    const int stride = kSpPipes * gridDim.x;                 // tiles of pipeline p: (3 * blockIdx.x + p) + i * stride

    if (kF64) {
        for (int i = tid; i < kNfft; i += kSpThreads) {
            sm.win[i] = (R)(2.0 * tb.win_half_d[i]);
            sm.w400[i] = mk<R>((R)tb.w400_d[i].x, (R)tb.w400_d[i].y);
        }
    } else {
        for (int i = tid; i < kNfft; i += kSpThreads) sm.win[i] = (R)(2.0f * tb.win_half[i]);
    }
    if (tid < kSpPipes) {
        SpPipe<R>& pp = sm.pipe[tid];
        for (int s = 0; s < kSpRawSlots; ++s) mbar_init(&pp.bar_raw_full[s], 32);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&pp.bar_span_full[b], 32);
            mbar_init(&pp.bar_span_empty[b], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    __syncthreads();

    if (tid >= kSpFftThreads) {
        // ============================== PREP warp of pipeline p ==============================
        // the prep warp group hands registers back, the three FFT warp groups take them (whole register file in use)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kSpRegPrep));
        const int p = (tid - kSpFftThreads) >> 5;
        if (p >= kSpPipes) return;                         // fourth warp of the group: idle
        const int lane = tid & 31;
        SpPipe<R>& pp = sm.pipe[p];
        const int first = kSpPipes * blockIdx.x + p;
        const double c = prm.pre_emphasis;
        auto stage = [&](int n, const WsTile& t) -> float {
            if (first + n * stride >= total_tiles) return 0.f;
            const int s = n % kSpRawSlots;
            if (lane == 0) pp.desc[n % kSpDescRing] = t;
            const float gain = __ldg(&stat[t.u].gain);
            const float* __restrict__ y = wav + t.sample_off;
            const int64_t q0 = (int64_t)t.t0 * kHop - kNfft / 2;
            float* dst = pp.raw[s];
            const float* __restrict__ src = y + (q0 - 4);
            if (!t.edge && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_expect_tx(&pp.bar_raw_full[s], kSpRaw * 4);
                    bulk_g2s(dst, src, kSpRaw * 4, &pp.bar_raw_full[s]);
                } else {
                    mbar_arrive(&pp.bar_raw_full[s]);
                }
            } else {
                if (!t.edge) {
                    for (int i = lane; i < kSpRaw; i += 32) cp_async4(dst + i, src + i);
                } else {
                    for (int i = lane; i < kSpRaw; i += 32) {
                        const int64_t q = q0 - 4 + i;
                        int64_t r = q < 0 ? -q : (q > t.L - 1 ? 2 * (t.L - 1) - q : q);
                        r = r < 0 ? 0 : (r > t.L - 1 ? t.L - 1 : r);
                        cp_async4(dst + i, y + r);
                    }
                }
                mbar_arrive_cp_async(&pp.bar_raw_full[s]);
            }
            __syncwarp();
            return gain;
        };
        auto fetch = [&](int n) -> WsTile {
            const int tile = first + n * stride;
            WsTile t;
            t.sample_off = 0; t.L = 1; t.frame_row = 0; t.u = 0; t.t0 = 0; t.nfr = 0; t.edge = 0;
            if (tile < total_tiles) t = tiles[tile];
            return t;
        };
        float g0 = stage(0, fetch(0));
        WsTile rec = fetch(1);
        for (int i = 0; first + i * stride < total_tiles; ++i) {
            // tile i+1 goes into the other raw slot (consumed by tile i-1) while tile i is converted
            const WsTile rec_next = fetch(i + 2);
            const float g1 = stage(i + 1, rec);
            const int s = i % kSpRawSlots, b = i & 1, k = i >> 1;
            mbar_wait(&pp.bar_raw_full[s], (i / kSpRawSlots) & 1);
            const WsTile t = pp.desc[i % kSpDescRing];
            const float gain = g0;
            mbar_wait(&pp.bar_span_empty[b], (k & 1) ^ 1);
            const float* __restrict__ raw = pp.raw[s] + 4;
            R* __restrict__ sp = pp.span[b];
            if (!t.edge) {
                constexpr int kBatch = 5;
                static_assert(kSpSpan % (64 * kBatch) == 0, "span must split into whole batches");
                for (int e0 = 2 * lane; e0 < kSpSpan; e0 += 64 * kBatch) {
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
                const int64_t q0 = (int64_t)t.t0 * kHop - kNfft / 2;
                constexpr int kBatchE = 8;
                static_assert(kSpSpan % (32 * kBatchE) == 0, "span must split into whole batches");
                for (int e0 = lane; e0 < kSpSpan; e0 += 32 * kBatchE) {
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
            mbar_arrive(&pp.bar_span_full[b]);
            g0 = g1;
            rec = rec_next;
        }
        return;
    }

    // ============================== FFT group of pipeline p ==============================
    {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kSpRegFft));
        const int p = tid >> 7;
        SpPipe<R>& pp = sm.pipe[p];
        const int first = kSpPipes * blockIdx.x + p;
        const int lt = tid & 127;
        const int lane = tid & 31;
        int ul1, j;
        bool s1_on = true;
        if (lt < 96) {
            ul1 = lt >> 4;
            j = lt & 15;
        } else {
            const int q = (lt - 96) >> 2;                 // 0..7 -> unit 0 2 1 3 4 - 5 -   (see k_fe_pass_a_ws)
            ul1 = q == 0 ? 0 : q == 1 ? 2 : q == 2 ? 1 : q == 3 ? 3 : q == 4 ? 4 : 5;
            s1_on = q != 5 && q != 7;
            j = 16 + (lt & 3);
        }
        // step-2 task: warps 0, 1 run 32 one-frame columns each, warps 2, 3 run 22 + the 6 packed columns c = 0 / 10;
        // pipelines alternate which pair is heavy so that the SM sub-partitions stay balanced
        const int lwl = ((lt >> 5) ^ ((p & 1) << 1)) & 3;
        int ul2 = 0, c2 = -1, row2 = 0;
        {
            int gt = -1;
            if (lwl < 2) gt = 32 * lwl + lane;
            else if (lane < 22) gt = 64 + 22 * (lwl - 2) + lane;
            else if (lane < 22 + kSpUnits) { ul2 = lane - 22; c2 = lwl == 2 ? 0 : 10; row2 = lwl == 2 ? 18 : 19; }
            if (gt >= 0) {
                ul2 = gt / 18;
                row2 = gt - ul2 * 18;
                c2 = row2 < 9 ? row2 + 1 : row2 + 2;
            }
        }
        cx<R> w1, w10;
        typename FeTw<R>::type tw;
        if (kF64) {
            w1 = sm.w400[j];
            w10 = sm.w400[10 * j];
        } else {
            tw.load(reinterpret_cast<const cx<R>*>(tb.w400), j);
        }
        cx<R>* slot_col = pp.slots + ul1 * kWsUnitSlots + j;
        const cx<R>* slot_row = pp.slots + ul2 * kWsUnitSlots + row2 * kSlotLd;
        float* power = reinterpret_cast<float*>(pp.slots);      // [12][201] staging tile, aliases the slots between barriers
        const int bar_id = 1 + p;
        for (int i = 0; first + i * stride < total_tiles; ++i) {
            const int b = i & 1, k = i >> 1;
            mbar_wait(&pp.bar_span_full[b], k & 1);
            if (s1_on) {
                const R* __restrict__ src = pp.span[b] + ul1 * (2 * kHop) + j;
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
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");      // A: slots written, span read
            mbar_arrive(&pp.bar_span_empty[b]);
            cx<R> v[20];
            if (c2 >= 0) {
#pragma unroll
                for (int n2 = 0; n2 < 20; ++n2) v[n2] = slot_row[n2];
            }
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");      // B: every slot row is in registers
            if (c2 >= 0) {
                dft20<false>(v);
                float* pa = power + (2 * ul2) * kBins;
                store_power(v, c2, pa, pa + kBins);
            }
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");      // C: the |X|^2 tile is complete
            // ---- coalesced store of |X|^2 (pass B takes the logarithm) + utterance max / min
            {
                const WsTile& t = pp.desc[i % kSpDescRing];            // still valid: the ring is 8 deep, prep stages at most 4 tiles ahead
                const int nfr = t.nfr;
                const int n = nfr * kBins;
                float* __restrict__ dst = pdb_out + t.frame_row * kBins;
                float p_max = 0.f, p_min = __int_as_float(0x7f800000);
                int head = 0;
                if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(power);
                    float4* __restrict__ d4 = reinterpret_cast<float4*>(dst);
                    const int n4 = n >> 2;
                    constexpr int kIt = (kSpFrames * kBins / 4 + 127) / 128;      // 5
                    float4 q[kIt];
#pragma unroll
                    for (int it = 0; it < kIt; ++it) {
                        const int e = lt + 128 * it;
                        if (e < n4) q[it] = p4[e];
                    }
#pragma unroll
                    for (int it = 0; it < kIt; ++it) {
                        const int e = lt + 128 * it;
                        if (e < n4) {
                            p_max = fmaxf(fmaxf(p_max, fmaxf(q[it].x, q[it].y)), fmaxf(q[it].z, q[it].w));
                            p_min = fminf(fminf(p_min, fminf(q[it].x, q[it].y)), fminf(q[it].z, q[it].w));
                            d4[e] = make_float4(pdb_store(q[it].x), pdb_store(q[it].y), pdb_store(q[it].z), pdb_store(q[it].w));
                        }
                    }
                    head = n4 << 2;
                }
                for (int e = head + lt; e < n; e += 128) {
                    const float x = power[e];
                    p_max = fmaxf(p_max, x);
                    p_min = fminf(p_min, x);
                    dst[e] = pdb_store(x);
                }
                const unsigned u_max = __reduce_max_sync(0xffffffffu, __float_as_uint(p_max));
                const unsigned u_min = __reduce_min_sync(0xffffffffu, __float_as_uint(p_min));
                if (lane == 0) {
                    atomicMax(&stat[t.u].p_max, u_max);
                    atomicMin(&stat[t.u].p_min, u_min);
                }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");      // D: staging read, slots free for the next tile
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Mel filterbank + raw mel dB + utterance max / min of the mel power from the stored |X|^2 (requires
// SC_DB_IN_PASS_B: the power-dB buffer holds |X|^2 between pass A and pass B).  One CTA = 32 frames of one utterance,
// 7 warps with the band ranges / pair records of k_fe_pass_a_ws, lane = frame.
static_assert(SC_DB_IN_PASS_B == 1, "k_fe_mel reads |X|^2 from the power-dB buffer");
constexpr int kMelFrames = 32;
constexpr int kMelThreads = 32 * kWsEpiWarps;       // 224

__global__ void __launch_bounds__(kMelThreads)
k_fe_mel(Ragged rg, UttStat* __restrict__ stat, const float* __restrict__ pdb, float* __restrict__ mel_raw,
         const int4* __restrict__ mel_brec, const float* __restrict__ mel_wt, WsMelParam mp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* P = reinterpret_cast<float*>(smem_raw);                               // [32][201] + 64 pad
    float* wt = P + kMelFrames * kBins + 64;                                      // n_taps
    int4* prec = reinterpret_cast<int4*>(wt + kWsMaxTaps);                        // kWsMaxPairs + 1
    float* ms = reinterpret_cast<float*>(prec + kWsMaxPairs + 1);                 // [32][n_mels | 1]
    const int tid = threadIdx.x, lane = tid & 31, ew = tid >> 5;
    const int n_mels = mp.n_mels, mel_ld = n_mels | 1;
    const int u = find_utt(rg.tile_prefix, rg.n_utts, blockIdx.x);
    const int t0 = (blockIdx.x - rg.tile_prefix[u]) * kMelFrames;
    const int nfr = min(kMelFrames, rg.frame_cnt[u] - t0);
    const int64_t row0 = rg.frame_off[u] + t0;
    // ---- stage the power tile (coalesced), the weights and the pair records
    {
        const float* __restrict__ src = pdb + row0 * kBins;
        const int n = nfr * kBins;
        int head = 0;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            const float4* __restrict__ s4 = reinterpret_cast<const float4*>(src);
            float4* __restrict__ d4 = reinterpret_cast<float4*>(P);
            const int n4 = n >> 2;
            constexpr int kIt = (kMelFrames * kBins / 4 + kMelThreads - 1) / kMelThreads;      // 8
            float4 q[kIt];
#pragma unroll
            for (int it = 0; it < kIt; ++it) {
                const int e = tid + kMelThreads * it;
                if (e < n4) q[it] = __ldg(s4 + e);
            }
#pragma unroll
            for (int it = 0; it < kIt; ++it) {
                const int e = tid + kMelThreads * it;
                if (e < n4) d4[e] = q[it];
            }
            head = n4 << 2;
        }
        for (int e = head + tid; e < n; e += kMelThreads) P[e] = __ldg(src + e);
        for (int e = n + tid; e < kMelFrames * kBins + 64; e += kMelThreads) P[e] = 0.f;      // missing frames, pad
        for (int i = tid; i < mp.n_taps; i += kMelThreads) wt[i] = mel_wt[i];
        for (int i = tid; i <= kWsMaxPairs; i += kMelThreads) prec[i] = mel_brec[i];
    }
    __syncthreads();
    // ---- sparse Slaney mel (:160-169), two bands per step (see k_fe_pass_a_ws)
    {
        const int mb = mp.chunk[ew], me = mp.chunk[ew + 1];
        const int pr0 = mp.pair0[ew], pr1 = mp.pair0[ew + 1];
        const float* __restrict__ prow = P + lane * kBins;
        float* __restrict__ mrow = ms + lane * mel_ld + mb;
        int4 d = prec[pr0];
        for (int pi = pr0; pi < pr1; ++pi) {
            const int4 nxt = prec[pi + 1];
            const float* __restrict__ pa = prow + d.x;
            const float* __restrict__ pb = prow + d.y;
            const float4* __restrict__ ww = reinterpret_cast<const float4*>(wt) + d.w;
            float acc_a = 0.f, acc_b = 0.f;
            int nb = d.z;
            while (nb > 4) {
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
    __syncthreads();
    // ---- raw amplitude_to_db (:172), coalesced rows, utterance max / min of the mel power
    {
        float m_max = 0.f, m_min = __int_as_float(0x7f800000);
        float* __restrict__ dst = mel_raw + row0 * n_mels;
        for (int f = ew; f < nfr; f += kWsEpiWarps)
            for (int m = lane; m < n_mels; m += 32) {
                const float v = ms[f * mel_ld + m];
                m_max = fmaxf(m_max, v);
                m_min = fminf(m_min, v);
                dst[f * n_mels + m] = 2.0f * db10(fmaxf(v, 1e-5f));
            }
        const unsigned u_max = __reduce_max_sync(0xffffffffu, __float_as_uint(m_max));
        const unsigned u_min = __reduce_min_sync(0xffffffffu, __float_as_uint(m_min));
        if (lane == 0) {
            atomicMax(&stat[u].m_max, u_max);
            atomicMin(&stat[u].m_min, u_min);
        }
    }
}

inline size_t fe_mel_smem_bytes(int n_mels) {
    return sizeof(float) * (kMelFrames * kBins + 64 + kWsMaxTaps) + sizeof(int4) * (kWsMaxPairs + 1) +
           sizeof(float) * kMelFrames * (n_mels | 1);
}

}  // namespace scdsp
