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

#define SC_HD __host__ __device__ __forceinline__

namespace scdsp {

constexpr int kNfft = 400;            // fast-path FFT length
constexpr int kBins = 201;            // 1 + kNfft/2
constexpr int kUnitThreads = 20;      // threads per unit (one frame pair)
constexpr int kSlotLd = 21;           // slot row stride (float2), odd => conflict-free both ways
constexpr int kUnitSlots = 20 * kSlotLd;  // 420 float2 per unit

SC_HD float sc_rsqrt(float x) {
#ifdef __CUDA_ARCH__
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}

// Per-thread twiddles: tw[k1] = W400^(j*k1) for k1 = 1..9, tw[0] = W400^(10*j).
struct Twiddle {
    float2 tw[10];
};

SC_HD void load_twiddles(Twiddle& t, const float2* __restrict__ w400, int j) {
#pragma unroll
    for (int k1 = 1; k1 < 10; ++k1) t.tw[k1] = w400[j * k1];
    t.tw[0] = w400[10 * j];
}

// ---- forward step 1: z[n1] = 0.5*w*(xA, xB) at sample 20*n1 + j  ->  slot column j
SC_HD void fwd_step1(float2 (&z)[20], const Twiddle& t, float2* __restrict__ slot_col) {
    dft20<false>(z);
    slot_col[0] = make_float2(z[0].x + z[0].x, z[0].y + z[0].y);
    const float2 z10 = make_float2(z[10].x + z[10].x, z[10].y + z[10].y);
    slot_col[10 * kSlotLd] = cmul(z10, t.tw[0]);
#pragma unroll
    for (int k1 = 1; k1 < 10; ++k1) {
        const float2 p = z[k1], q = z[20 - k1];
        const float2 ya = make_float2(p.x + q.x, p.y - q.y);
        const float2 yb = make_float2(p.y + q.y, q.x - p.x);
        slot_col[k1 * kSlotLd] = cmul(ya, t.tw[k1]);
        slot_col[(10 + k1) * kSlotLd] = cmul(yb, t.tw[k1]);
    }
}

// ---- forward step 2: slot row c -> V[k2]
SC_HD void fwd_step2(float2 (&v)[20], const float2* __restrict__ slot_row) {
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) v[n2] = slot_row[n2];
    dft20<false>(v);
}

// bin owned by generic column (k1 in 1..9) at position k2
SC_HD constexpr int own_bin(int k1, int k2) { return k2 < 10 ? k1 + 20 * k2 : (20 - k1) + 20 * (19 - k2); }

// ---- |X|^2 of the bins owned by column thread c, written to the two frames' power rows
//      (audio_lib.py:150-155: F = |stft|, P = F**2)
SC_HD void store_power(const float2 (&v)[20], int c, float* __restrict__ pa, float* __restrict__ pb) {
    if (c == 0) {
#pragma unroll
        for (int k2 = 0; k2 <= 10; ++k2) {
            const float2 p = v[k2], q = v[(20 - k2) % 20];
            const float ar = p.x + q.x, ai = p.y - q.y, br = p.y + q.y, bi = q.x - p.x;
            pa[20 * k2] = 0.25f * fmaf(ar, ar, ai * ai);
            pb[20 * k2] = 0.25f * fmaf(br, br, bi * bi);
        }
    } else if (c == 10) {
#pragma unroll
        for (int k2 = 0; k2 < 10; ++k2) {
            const float2 p = v[k2], q = v[19 - k2];
            const float ar = p.x + q.x, ai = p.y - q.y, br = p.y + q.y, bi = q.x - p.x;
            pa[10 + 20 * k2] = 0.25f * fmaf(ar, ar, ai * ai);
            pb[10 + 20 * k2] = 0.25f * fmaf(br, br, bi * bi);
        }
    } else {
        const int k1 = c < 10 ? c : c - 10;
        float* __restrict__ row = c < 10 ? pa : pb;
#pragma unroll
        for (int k2 = 0; k2 < 10; ++k2) row[k1 + 20 * k2] = fmaf(v[k2].x, v[k2].x, v[k2].y * v[k2].y);
#pragma unroll
        for (int k2 = 10; k2 < 20; ++k2)
            row[(20 - k1) + 20 * (19 - k2)] = fmaf(v[k2].x, v[k2].x, v[k2].y * v[k2].y);
    }
}

// unit-phase * amplitude (audio_lib.py:268-270: S = A * exp(1j*angle(X)); angle(0) = 0)
SC_HD float2 impose(float2 x, float a) {
    const float n2 = fmaf(x.x, x.x, x.y * x.y);
    if (n2 > 0.0f) {
        const float s = a * sc_rsqrt(n2);
        return make_float2(x.x * s, x.y * s);
    }
    return make_float2(a, 0.0f);
}

// ---- Griffin-Lim phase update on the bins owned by column thread c, leaving v ready for the
//      inverse step 2'.  amp_a / amp_b are the two frames' magnitude rows (201 floats each).
SC_HD void gl_update(float2 (&v)[20], int c, const float* __restrict__ amp_a, const float* __restrict__ amp_b) {
    if (c == 0) {
#pragma unroll
        for (int k2 = 0; k2 <= 10; ++k2) {
            const int m = (20 - k2) % 20;
            const float2 p = v[k2], q = v[m];
            const float2 sa = impose(make_float2(p.x + q.x, p.y - q.y), amp_a[20 * k2]);
            const float2 sb = impose(make_float2(p.y + q.y, q.x - p.x), amp_b[20 * k2]);
            v[k2] = make_float2(sa.x - sb.y, sa.y + sb.x);
            if (m != k2) v[m] = make_float2(sa.x + sb.y, sb.x - sa.y);
        }
    } else if (c == 10) {
#pragma unroll
        for (int k2 = 0; k2 < 10; ++k2) {
            const int m = 19 - k2;
            const float2 p = v[k2], q = v[m];
            const float2 sa = impose(make_float2(p.x + q.x, p.y - q.y), amp_a[10 + 20 * k2]);
            const float2 sb = impose(make_float2(p.y + q.y, q.x - p.x), amp_b[10 + 20 * k2]);
            v[k2] = make_float2(sa.x - sb.y, sa.y + sb.x);
            v[m] = make_float2(sa.x + sb.y, sb.x - sa.y);
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
SC_HD float2 polar(float a, float ph) {
    float s, c;
#ifdef __CUDA_ARCH__
    sincosf(ph, &s, &c);
#else
    s = sinf(ph); c = cosf(ph);
#endif
    return make_float2(a * c, a * s);
}

SC_HD void gl_init_state(float2 (&v)[20], int c, const float* __restrict__ amp_a, const float* __restrict__ amp_b,
                         const float* __restrict__ ph_a, const float* __restrict__ ph_b) {
    if (c == 0) {
#pragma unroll
        for (int k2 = 0; k2 <= 10; ++k2) {
            const int m = (20 - k2) % 20;
            float2 sa = polar(amp_a[20 * k2], ph_a[20 * k2]);
            float2 sb = polar(amp_b[20 * k2], ph_b[20 * k2]);
            if (k2 == 0 || k2 == 10) { sa.y = 0.0f; sb.y = 0.0f; }
            v[k2] = make_float2(sa.x - sb.y, sa.y + sb.x);
            if (m != k2) v[m] = make_float2(sa.x + sb.y, sb.x - sa.y);
        }
    } else if (c == 10) {
#pragma unroll
        for (int k2 = 0; k2 < 10; ++k2) {
            const int m = 19 - k2;
            const float2 sa = polar(amp_a[10 + 20 * k2], ph_a[10 + 20 * k2]);
            const float2 sb = polar(amp_b[10 + 20 * k2], ph_b[10 + 20 * k2]);
            v[k2] = make_float2(sa.x - sb.y, sa.y + sb.x);
            v[m] = make_float2(sa.x + sb.y, sb.x - sa.y);
        }
    } else {
        const int k1 = c < 10 ? c : c - 10;
        const float* __restrict__ amp = c < 10 ? amp_a : amp_b;
        const float* __restrict__ ph = c < 10 ? ph_a : ph_b;
#pragma unroll
        for (int k2 = 0; k2 < 20; ++k2) {
            const int b = own_bin(k1, k2);
            float2 s = polar(amp[b], ph[b]);
            if (k2 >= 10) s.y = -s.y;
            v[k2] = s;
        }
    }
}

// ---- inverse step 2': u[k2] = S[c-column] -> 20-point inverse DFT over k2 -> slot row c
SC_HD void inv_step2(float2 (&u)[20], float2* __restrict__ slot_row) {
    dft20<true>(u);
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) slot_row[n2] = u[n2];
}

// ---- inverse step 1': slot column j -> h[k1] -> 20-point inverse DFT over k1.
//      On return h[n1] = 400 * (xA[20*n1 + j], xB[20*n1 + j]).
SC_HD void inv_step1(float2 (&h)[20], const Twiddle& t, const float2* __restrict__ slot_col) {
    h[0] = slot_col[0];
    h[10] = cmulc(slot_col[10 * kSlotLd], t.tw[0]);
#pragma unroll
    for (int k1 = 1; k1 < 10; ++k1) {
        const float2 ha = cmulc(slot_col[k1 * kSlotLd], t.tw[k1]);
        const float2 hb = cmulc(slot_col[(10 + k1) * kSlotLd], t.tw[k1]);
        h[k1] = make_float2(ha.x - hb.y, ha.y + hb.x);
        h[20 - k1] = make_float2(ha.x + hb.y, hb.x - ha.y);
    }
    dft20<true>(h);
}

}  // namespace scdsp
