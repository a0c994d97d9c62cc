// 20-point complex DFT held entirely in registers (Good-Thomas 4 x 5, no internal twiddles).
//
// Building block of the real FFT-400 used by both halves of the hot path:
//   STFT   of audio_lib.py:141-147 / :267   (librosa.stft  -> scipy FFT of n_fft = 400)
//   iSTFT  of audio_lib.py:260              (librosa.istft -> scipy inverse FFT)
// n_fft = 400 = 20 x 20 in every shipped hp/*.json (win_length_ms 25 @ 16 kHz).
//
// Templated on the real type R: the front-end runs it in float64 by default (the reference's FFT
// is float64 and bins 70-80 dB below the utterance maximum are only reproduced to the 1e-5
// tolerance with more than 24 mantissa bits; B200 issues FP64 FMAs at half the FP32 rate),
// Griffin-Lim and the opt-in fast front-end run it in float32.
#pragma once
#include <cuda_runtime.h>

#define SC_HD __host__ __device__ __forceinline__

namespace scdsp {

template <typename R> struct alignas(2 * sizeof(R)) cx { R x, y; };
using cxf = cx<float>;
using cxd = cx<double>;

template <typename R> SC_HD cx<R> mk(R x, R y) { cx<R> c; c.x = x; c.y = y; return c; }
template <typename R> SC_HD cx<R> cadd(cx<R> a, cx<R> b) { return mk<R>(a.x + b.x, a.y + b.y); }
template <typename R> SC_HD cx<R> csub(cx<R> a, cx<R> b) { return mk<R>(a.x - b.x, a.y - b.y); }
SC_HD float sc_fma(float a, float b, float c) { return fmaf(a, b, c); }
SC_HD double sc_fma(double a, double b, double c) { return fma(a, b, c); }
// a * b
template <typename R> SC_HD cx<R> cmul(cx<R> a, cx<R> b) {
    return mk<R>(sc_fma(a.x, b.x, -a.y * b.y), sc_fma(a.x, b.y, a.y * b.x));
}
// a * conj(b)
template <typename R> SC_HD cx<R> cmulc(cx<R> a, cx<R> b) {
    return mk<R>(sc_fma(a.x, b.x, a.y * b.y), sc_fma(a.y, b.x, -a.x * b.y));
}

// 4-point DFT, in place.  INV = false: kernel exp(-2*pi*i*n*k/4); INV = true: exp(+...).
template <bool INV, typename R>
SC_HD void radix4(cx<R>& a0, cx<R>& a1, cx<R>& a2, cx<R>& a3) {
    const cx<R> s0 = cadd(a0, a2), d0 = csub(a0, a2);
    const cx<R> s1 = cadd(a1, a3), d1 = csub(a1, a3);
    a0 = cadd(s0, s1);
    a2 = csub(s0, s1);
    // forward: X1 = d0 - i*d1, X3 = d0 + i*d1
    const cx<R> p = mk<R>(d0.x + d1.y, d0.y - d1.x);
    const cx<R> m = mk<R>(d0.x - d1.y, d0.y + d1.x);
    a1 = INV ? m : p;
    a3 = INV ? p : m;
}

// 5-point DFT, in place.
template <bool INV, typename R>
SC_HD void radix5(cx<R>& a0, cx<R>& a1, cx<R>& a2, cx<R>& a3, cx<R>& a4) {
    constexpr R C1 = (R)0.30901699437494742410;    // cos(2*pi/5)
    constexpr R C2 = (R)-0.80901699437494742410;   // cos(4*pi/5)
    constexpr R S1 = (R)0.95105651629515357212;    // sin(2*pi/5)
    constexpr R S2 = (R)0.58778525229247312917;    // sin(4*pi/5)
    const cx<R> t1 = cadd(a1, a4), t3 = csub(a1, a4);
    const cx<R> t2 = cadd(a2, a3), t4 = csub(a2, a3);
    const cx<R> m1 = mk<R>(sc_fma(C2, t2.x, sc_fma(C1, t1.x, a0.x)), sc_fma(C2, t2.y, sc_fma(C1, t1.y, a0.y)));
    const cx<R> m2 = mk<R>(sc_fma(C1, t2.x, sc_fma(C2, t1.x, a0.x)), sc_fma(C1, t2.y, sc_fma(C2, t1.y, a0.y)));
    const cx<R> q1 = mk<R>(sc_fma(S2, t4.x, S1 * t3.x), sc_fma(S2, t4.y, S1 * t3.y));
    const cx<R> q2 = mk<R>(sc_fma(-S1, t4.x, S2 * t3.x), sc_fma(-S1, t4.y, S2 * t3.y));
    a0 = mk<R>(a0.x + t1.x + t2.x, a0.y + t1.y + t2.y);
    // forward: X1 = m1 - i*q1, X4 = m1 + i*q1, X2 = m2 - i*q2, X3 = m2 + i*q2
    const cx<R> x1 = mk<R>(m1.x + q1.y, m1.y - q1.x);
    const cx<R> x4 = mk<R>(m1.x - q1.y, m1.y + q1.x);
    const cx<R> x2 = mk<R>(m2.x + q2.y, m2.y - q2.x);
    const cx<R> x3 = mk<R>(m2.x - q2.y, m2.y + q2.x);
    a1 = INV ? x4 : x1;
    a4 = INV ? x1 : x4;
    a2 = INV ? x3 : x2;
    a3 = INV ? x2 : x3;
}

// Unnormalised 20-point DFT, natural order in and out, fully unrolled (register resident).
//   input  index n = (5*n1 + 4*n2) mod 20,  n1 in [0,4), n2 in [0,5)
//   output index k = (5*k1 + 16*k2) mod 20
template <bool INV, typename R>
SC_HD void dft20(cx<R> (&v)[20]) {
    cx<R> t[4][5];
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) {
        cx<R> a0 = v[(0 + 4 * n2) % 20], a1 = v[(5 + 4 * n2) % 20];
        cx<R> a2 = v[(10 + 4 * n2) % 20], a3 = v[(15 + 4 * n2) % 20];
        radix4<INV>(a0, a1, a2, a3);
        t[0][n2] = a0; t[1][n2] = a1; t[2][n2] = a2; t[3][n2] = a3;
    }
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        radix5<INV>(t[k1][0], t[k1][1], t[k1][2], t[k1][3], t[k1][4]);
#pragma unroll
        for (int k2 = 0; k2 < 5; ++k2) v[(5 * k1 + 16 * k2) % 20] = t[k1][k2];
    }
}

}  // namespace scdsp
