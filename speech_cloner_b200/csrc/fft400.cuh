// Real FFT-400 / inverse real FFT-400 for a PAIR of frames, computed by a "unit" of 20 threads.
//
//   n = 20*n1 + n2,  k = k1 + 20*k2   (n1, n2, k1, k2 in [0, 20))
//   X[k1 + 20*k2] = sum_n2 W20^(n2*k2) * [ W400^(n2*k1) * sum_n1 x[20*n1 + n2] * W20^(n1*k1) ]
//
// step 1  thread j (= n2) packs the two real frames as z = xA + i*xB, runs one 20-point DFT over n1,
//         splits it into the Hermitian halves YA[k1], YB[k1] (k1 = 0..10), applies W400^(j*k1)
//         and writes 20 complex "slots" (row = column task c, column = j) to shared memory:
//           c = 0       (YA[0],  YB[0])                  both real, packed as one complex
//           c = 1..9    YA[c]  * W400^(j*c)              frame A
//           c = 10      (YA[10] + i*YB[10]) * W400^(10*j) both real before the twiddle
//           c = 11..19  YB[c-10] * W400^(j*(c-10))       frame B
// step 2  thread c runs one 20-point DFT over n2 of slot row c:
//           c = 1..9 / 11..19: V[k2] = X[k1 + 20*k2] of one frame; k2 >= 10 are bins > 200, i.e.
//                              the conjugates of bins (20-k1) + 20*(19-k2) <= 200
//           c = 0:  V = XA[20*k2] + i*XB[20*k2]          split with V[(20-k2)%20]
//           c = 10: V = XA[10+20*k2] + i*XB[10+20*k2]    split with V[19-k2]
//         so every thread ends up owning ~20 of the 2 x 201 bins, with no mirror exchange.
// The inverse runs the same graph backwards (steps 2' and 1').
//
// Replaces the FFTs inside librosa.stft / librosa.istft as called at audio_lib.py:141-147,
// :260 and :267.  All routines are __host__ __device__ so tests/host can run the exact index
// math on the CPU (one loop iteration per emulated thread).
#pragma once
#include "dft20.cuh"

namespace scdsp {

constexpr int kNfft = 400;            // fast-path FFT length
constexpr int kBins = 201;            // 1 + kNfft/2
constexpr int kUnitThreads = 20;      // threads per unit (one frame pair)
constexpr int kSlotLd = 21;           // slot row stride (complex), odd => conflict-free both ways
constexpr int kUnitSlots = 20 * kSlotLd;  // 420 complex per unit

SC_HD float sc_rsqrt(float x) {
#ifdef __CUDA_ARCH__
    float r;                                   // one MUFU.RSQ; rsqrtf() adds ~8 instructions of denormal
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));   // handling that the phase update does not need
    return r;
#else
    return 1.0f / sqrtf(x);
#endif
}

// Per-thread twiddles W400^(j*k1), k1 = 1..9, and W400^(10*j).
// TwReg keeps them in registers (float32 kernels); TwTab reads a (shared-memory) table on use
// (float64 front-end, where 19 complex doubles would cost 76 registers).
template <typename R> struct TwReg {
    cx<R> tw[10];
    SC_HD void load(const cx<R>* __restrict__ w400, int j) {
#pragma unroll
        for (int k1 = 1; k1 < 10; ++k1) tw[k1] = w400[j * k1];
        tw[0] = w400[10 * j];
    }
    SC_HD cx<R> get(int k1) const { return tw[k1]; }
    SC_HD cx<R> get10() const { return tw[0]; }
};
template <typename R> struct TwTab {
    const cx<R>* tab;
    int j;
    SC_HD void load(const cx<R>* w400, int j_) { tab = w400; j = j_; }
    SC_HD cx<R> get(int k1) const { return tab[j * k1]; }
    SC_HD cx<R> get10() const { return tab[10 * j]; }
};
using Twiddle = TwReg<float>;
SC_HD void load_twiddles(Twiddle& t, const cxf* __restrict__ w400, int j) { t.load(w400, j); }

// ---- forward step 1: z[n1] = 0.5*w*(xA, xB) at sample 20*n1 + j  ->  slot column j
template <typename R, typename TW>
SC_HD void fwd_step1(cx<R> (&z)[20], const TW& t, cx<R>* __restrict__ slot_col) {
    dft20<false>(z);
    slot_col[0] = mk<R>(z[0].x + z[0].x, z[0].y + z[0].y);
    const cx<R> z10 = mk<R>(z[10].x + z[10].x, z[10].y + z[10].y);
    slot_col[10 * kSlotLd] = cmul(z10, t.get10());
#pragma unroll
    for (int k1 = 1; k1 < 10; ++k1) {
        const cx<R> p = z[k1], q = z[20 - k1];
        const cx<R> ya = mk<R>(p.x + q.x, p.y - q.y);
        const cx<R> yb = mk<R>(p.y + q.y, q.x - p.x);
        const cx<R> w = t.get(k1);
        slot_col[k1 * kSlotLd] = cmul(ya, w);
        slot_col[(10 + k1) * kSlotLd] = cmul(yb, w);
    }
}

// ---- forward step 2: slot row c -> V[k2]
template <typename R>
SC_HD void fwd_step2(cx<R> (&v)[20], const cx<R>* __restrict__ slot_row) {
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) v[n2] = slot_row[n2];
    dft20<false>(v);
}

// Step-2 work assignment.  A step-2 task is (unit, column c); its inputs come from shared memory, so any
// thread can run any task.  The packed columns c = 0 and c = 10 take a different code path than the
// 18 one-frame columns: park them on the first lanes of the first two warps so that all other warps run
// a single path instead of every warp executing all three.
template <bool SPLIT>
SC_HD void step2_task(int tid, int units, int& unit, int& c) {
    if (SPLIT) {
        // c = 0 tasks on the first lanes of warp 0, c = 10 tasks on the first lanes of warp 1 (units <= 16): each
        // of those two warps runs ONE packed path next to the generic one (measured best for the front-end)
        if (tid < units) {
            unit = tid; c = 0;
            return;
        }
        if (tid >= 32 && tid < 32 + units) {
            unit = tid - 32; c = 10;
            return;
        }
        const int g = tid < 32 ? tid - units : tid - 2 * units;
        unit = g / 18;
        const int r = g - unit * 18;
        c = r < 9 ? r + 1 : r + 2;
    } else {
        // all packed columns on the lowest thread ids: with 16 units warp 0 is entirely packed columns and
        // warps 1..9 run only the generic path (measured best for Griffin-Lim)
        if (tid < 2 * units) {
            unit = tid >> 1;
            c = (tid & 1) ? 10 : 0;
            return;
        }
        const int g = tid - 2 * units;
        unit = g / 18;
        const int r = g - unit * 18;
        c = r < 9 ? r + 1 : r + 2;
    }
}

// bin owned by generic column (k1 in 1..9) at position k2
SC_HD constexpr int own_bin(int k1, int k2) { return k2 < 10 ? k1 + 20 * k2 : (20 - k1) + 20 * (19 - k2); }

// ---- |X|^2 of the bins owned by column thread c, written to the two frames' power rows
//      (audio_lib.py:150-155: F = |stft|, P = F**2)
template <typename R>
SC_HD void store_power(const cx<R> (&v)[20], int c, float* __restrict__ pa, float* __restrict__ pb) {
    if (c == 0) {
#pragma unroll
        for (int k2 = 0; k2 <= 10; ++k2) {
            const cx<R> p = v[k2], q = v[(20 - k2) % 20];
            const R ar = p.x + q.x, ai = p.y - q.y, br = p.y + q.y, bi = q.x - p.x;
            pa[20 * k2] = (float)((R)0.25 * sc_fma(ar, ar, ai * ai));
            pb[20 * k2] = (float)((R)0.25 * sc_fma(br, br, bi * bi));
        }
    } else if (c == 10) {
#pragma unroll
        for (int k2 = 0; k2 < 10; ++k2) {
            const cx<R> p = v[k2], q = v[19 - k2];
            const R ar = p.x + q.x, ai = p.y - q.y, br = p.y + q.y, bi = q.x - p.x;
            pa[10 + 20 * k2] = (float)((R)0.25 * sc_fma(ar, ar, ai * ai));
            pb[10 + 20 * k2] = (float)((R)0.25 * sc_fma(br, br, bi * bi));
        }
    } else {
        const int k1 = c < 10 ? c : c - 10;
        float* __restrict__ row = c < 10 ? pa : pb;
#pragma unroll
        for (int k2 = 0; k2 < 10; ++k2) row[k1 + 20 * k2] = (float)sc_fma(v[k2].x, v[k2].x, v[k2].y * v[k2].y);
#pragma unroll
        for (int k2 = 10; k2 < 20; ++k2)
            row[(20 - k1) + 20 * (19 - k2)] = (float)sc_fma(v[k2].x, v[k2].x, v[k2].y * v[k2].y);
    }
}

// unit-phase * amplitude (audio_lib.py:268-270: S = A * exp(1j*angle(X)); angle(0) = 0)
SC_HD cxf impose(cxf x, float a) {
    const float n2 = fmaf(x.x, x.x, x.y * x.y);
    if (n2 > 0.0f) {
        const float s = a * sc_rsqrt(n2);
        return mk<float>(x.x * s, x.y * s);
    }
    return mk<float>(a, 0.0f);
}

// ---- Griffin-Lim phase update on the bins owned by column thread c, leaving v ready for the
//      inverse step 2'.  amp_a / amp_b are the two frames' magnitude rows (201 floats each).
SC_HD void gl_update(cxf (&v)[20], int c, const float* __restrict__ amp_a, const float* __restrict__ amp_b) {
    if (c == 0) {
#pragma unroll
        for (int k2 = 0; k2 <= 10; ++k2) {
            const int m = (20 - k2) % 20;
            const cxf p = v[k2], q = v[m];
            const cxf sa = impose(mk<float>(p.x + q.x, p.y - q.y), amp_a[20 * k2]);
            const cxf sb = impose(mk<float>(p.y + q.y, q.x - p.x), amp_b[20 * k2]);
            v[k2] = mk<float>(sa.x - sb.y, sa.y + sb.x);
            if (m != k2) v[m] = mk<float>(sa.x + sb.y, sb.x - sa.y);
        }
    } else if (c == 10) {
#pragma unroll
        for (int k2 = 0; k2 < 10; ++k2) {
            const int m = 19 - k2;
            const cxf p = v[k2], q = v[m];
            const cxf sa = impose(mk<float>(p.x + q.x, p.y - q.y), amp_a[10 + 20 * k2]);
            const cxf sb = impose(mk<float>(p.y + q.y, q.x - p.x), amp_b[10 + 20 * k2]);
            v[k2] = mk<float>(sa.x - sb.y, sa.y + sb.x);
            v[m] = mk<float>(sa.x + sb.y, sb.x - sa.y);
        }
    } else {
        const int k1 = c < 10 ? c : c - 10;
        const float* __restrict__ amp = c < 10 ? amp_a : amp_b;
#pragma unroll
        for (int k2 = 0; k2 < 20; ++k2) v[k2] = impose(v[k2], amp[own_bin(k1, k2)]);
    }
}

// ---- Initial Griffin-Lim state S0 = A * exp(i*phase0) (audio_lib.py:255-256) laid out as the
//      input of inverse step 2'.  The imaginary parts of bins 0 and 200 are dropped, as the
//      reference's Hermitian extension + ".real" does inside librosa.istft.
SC_HD cxf polar(float a, float ph) {
    float s, c;
#ifdef __CUDA_ARCH__
    sincosf(ph, &s, &c);
#else
    s = sinf(ph); c = cosf(ph);
#endif
    return mk<float>(a * c, a * s);
}

SC_HD void gl_init_state(cxf (&v)[20], int c, const float* __restrict__ amp_a, const float* __restrict__ amp_b,
                         const float* __restrict__ ph_a, const float* __restrict__ ph_b) {
    if (c == 0) {
#pragma unroll
        for (int k2 = 0; k2 <= 10; ++k2) {
            const int m = (20 - k2) % 20;
            cxf sa = polar(amp_a[20 * k2], ph_a[20 * k2]);
            cxf sb = polar(amp_b[20 * k2], ph_b[20 * k2]);
            if (k2 == 0 || k2 == 10) { sa.y = 0.0f; sb.y = 0.0f; }
            v[k2] = mk<float>(sa.x - sb.y, sa.y + sb.x);
            if (m != k2) v[m] = mk<float>(sa.x + sb.y, sb.x - sa.y);
        }
    } else if (c == 10) {
#pragma unroll
        for (int k2 = 0; k2 < 10; ++k2) {
            const int m = 19 - k2;
            const cxf sa = polar(amp_a[10 + 20 * k2], ph_a[10 + 20 * k2]);
            const cxf sb = polar(amp_b[10 + 20 * k2], ph_b[10 + 20 * k2]);
            v[k2] = mk<float>(sa.x - sb.y, sa.y + sb.x);
            v[m] = mk<float>(sa.x + sb.y, sb.x - sa.y);
        }
    } else {
        const int k1 = c < 10 ? c : c - 10;
        const float* __restrict__ amp = c < 10 ? amp_a : amp_b;
        const float* __restrict__ ph = c < 10 ? ph_a : ph_b;
#pragma unroll
        for (int k2 = 0; k2 < 20; ++k2) {
            const int b = own_bin(k1, k2);
            cxf s = polar(amp[b], ph[b]);
            if (k2 >= 10) s.y = -s.y;
            v[k2] = s;
        }
    }
}

// ---- inverse step 2': u[k2] = S[c-column] -> 20-point inverse DFT over k2 -> slot row c
template <typename R>
SC_HD void inv_step2(cx<R> (&u)[20], cx<R>* __restrict__ slot_row) {
    dft20<true>(u);
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) slot_row[n2] = u[n2];
}

// ---- inverse step 1': slot column j -> h[k1] -> 20-point inverse DFT over k1.
//      On return h[n1] = 400 * (xA[20*n1 + j], xB[20*n1 + j]).
template <typename R, typename TW>
SC_HD void inv_step1(cx<R> (&h)[20], const TW& t, const cx<R>* __restrict__ slot_col) {
    h[0] = slot_col[0];
    h[10] = cmulc(slot_col[10 * kSlotLd], t.get10());
#pragma unroll
    for (int k1 = 1; k1 < 10; ++k1) {
        const cx<R> w = t.get(k1);
        const cx<R> ha = cmulc(slot_col[k1 * kSlotLd], w);
        const cx<R> hb = cmulc(slot_col[(10 + k1) * kSlotLd], w);
        h[k1] = mk<R>(ha.x - hb.y, ha.y + hb.x);
        h[20 - k1] = mk<R>(ha.x + hb.y, hb.x - ha.y);
    }
    dft20<true>(h);
}

// ---- real-input variants of step 1 (forward) and step 1' (inverse): two real 20-point DFTs instead of one packed
//      complex one + Hermitian split.  xa / xb are the FULL-window samples of frames A / B at 20*n1 + j.
template <typename R, typename TW>
SC_HD void fwd_step1_real(const R (&xa)[20], const R (&xb)[20], const TW& t, cx<R>* __restrict__ slot_col) {
    cx<R> ya[11], yb[11];
    rdft20_fwd(xa, ya);
    rdft20_fwd(xb, yb);
    slot_col[0] = mk<R>(ya[0].x, yb[0].x);
    slot_col[10 * kSlotLd] = cmul(mk<R>(ya[10].x, yb[10].x), t.get10());
#pragma unroll
    for (int k1 = 1; k1 < 10; ++k1) {
        const cx<R> w = t.get(k1);
        slot_col[k1 * kSlotLd] = cmul(ya[k1], w);
        slot_col[(10 + k1) * kSlotLd] = cmul(yb[k1], w);
    }
}

// On return xa[n1] = 400 * xA[20*n1 + j], xb[n1] = 400 * xB[20*n1 + j].
template <typename R, typename TW>
SC_HD void inv_step1_real(R (&xa)[20], R (&xb)[20], const TW& t, const cx<R>* __restrict__ slot_col) {
    cx<R> ha[11], hb[11];
    const cx<R> s0 = slot_col[0];
    const cx<R> s10 = cmulc(slot_col[10 * kSlotLd], t.get10());
    ha[0] = mk<R>(s0.x, (R)0); hb[0] = mk<R>(s0.y, (R)0);
    ha[10] = mk<R>(s10.x, (R)0); hb[10] = mk<R>(s10.y, (R)0);
#pragma unroll
    for (int k1 = 1; k1 < 10; ++k1) {
        const cx<R> w = t.get(k1);
        ha[k1] = cmulc(slot_col[k1 * kSlotLd], w);
        hb[k1] = cmulc(slot_col[(10 + k1) * kSlotLd], w);
    }
    rdft20_inv(ha, xa);
    rdft20_inv(hb, xb);
}

}  // namespace scdsp
