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
SC_HD void dft20_scalar(cx<R> (&v)[20]) {
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

template <bool INV, typename R>
SC_HD void dft20(cx<R> (&v)[20]) {
    dft20_scalar<INV>(v);
}
// (A packed FADD2 / FFMA2 form of these butterflies was measured in round 1 and was slower under Griffin-Lim's
// 96-register cap: 0.0758 -> 0.0816 ms per iteration; it is not part of the product.)

// ---------------------------------------------------------------------------------------------------
// Real-input 20-point DFT (forward) and its Hermitian-input inverse, same Good-Thomas 4 x 5 graph with
// the conjugate-symmetric half of the work removed (~94 instead of 224 instructions).
//   forward:  Y[k] = sum_n x[n] exp(-2*pi*i*n*k/20),            k = 0..10  (Y[20-k] = conj Y[k])
//   inverse:  x[n] = sum_{k=0..19} Y[k] exp(+2*pi*i*n*k/20)     with the Hermitian extension of Y[0..10];
//             the imaginary parts of Y[0] and Y[10] are ignored.

// 5-point DFT of real inputs: z0 real, z1 = (m1, -q1), z2 = (m2, -q2)  (z3 = conj z2, z4 = conj z1)
template <typename R>
SC_HD void rradix5(R a0, R a1, R a2, R a3, R a4, R& z0, cx<R>& z1, cx<R>& z2) {
    constexpr R C1 = (R)0.30901699437494742410, C2 = (R)-0.80901699437494742410;
    constexpr R S1 = (R)0.95105651629515357212, S2 = (R)0.58778525229247312917;
    const R t1 = a1 + a4, t3 = a1 - a4, t2 = a2 + a3, t4 = a2 - a3;
    z0 = a0 + t1 + t2;
    z1 = mk<R>(sc_fma(C2, t2, sc_fma(C1, t1, a0)), -sc_fma(S2, t4, S1 * t3));
    z2 = mk<R>(sc_fma(C1, t2, sc_fma(C2, t1, a0)), -sc_fma(-S1, t4, S2 * t3));
}
// inverse of rradix5: real outputs t[n] = z0 + 2 Re(z1 e^{+i n th}) + 2 Re(z2 e^{+2 i n th}), th = 2*pi/5
template <typename R>
SC_HD void rradix5_inv(R z0, cx<R> z1, cx<R> z2, R& t0, R& t1, R& t2, R& t3, R& t4) {
    constexpr R C1 = (R)0.30901699437494742410, C2 = (R)-0.80901699437494742410;
    constexpr R S1 = (R)0.95105651629515357212, S2 = (R)0.58778525229247312917;
    const R u1 = z1.x + z1.x, u2 = z2.x + z2.x, w1 = z1.y + z1.y, w2 = z2.y + z2.y;
    const R a = sc_fma(C2, u2, sc_fma(C1, u1, z0)), b = sc_fma(C1, u2, sc_fma(C2, u1, z0));
    const R e = sc_fma(S2, w2, S1 * w1), f = sc_fma(-S1, w2, S2 * w1);
    t0 = z0 + u1 + u2;
    t1 = a - e; t4 = a + e;
    t2 = b - f; t3 = b + f;
}

template <typename R>
SC_HD void rdft20_fwd(const R (&x)[20], cx<R> (&Y)[11]) {
    R t0[5], t2[5];
    cx<R> t1[5];
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) {
        const R a0 = x[(0 + 4 * n2) % 20], a1 = x[(5 + 4 * n2) % 20], a2 = x[(10 + 4 * n2) % 20], a3 = x[(15 + 4 * n2) % 20];
        const R r0 = a0 + a2, r1 = a0 - a2, r2 = a1 + a3, r3 = a1 - a3;
        t0[n2] = r0 + r2;
        t2[n2] = r0 - r2;
        t1[n2] = mk<R>(r1, -r3);
    }
    R z0; cx<R> z1, z2;
    // k1 = 0: k = 16*k2 mod 20 -> 0, 16, 12 (8 = conj 12, 4 = conj 16)
    rradix5(t0[0], t0[1], t0[2], t0[3], t0[4], z0, z1, z2);
    Y[0] = mk<R>(z0, (R)0);
    Y[4] = mk<R>(z1.x, -z1.y);
    Y[8] = mk<R>(z2.x, -z2.y);
    // k1 = 2: k = 10 + 16*k2 mod 20 -> 10, 6, 2
    rradix5(t2[0], t2[1], t2[2], t2[3], t2[4], z0, z1, z2);
    Y[10] = mk<R>(z0, (R)0);
    Y[6] = z1;
    Y[2] = z2;
    // k1 = 1: k = 5 + 16*k2 mod 20 -> 5, 1, 17, 13, 9 (3 = conj 17, 7 = conj 13)
    radix5<false>(t1[0], t1[1], t1[2], t1[3], t1[4]);
    Y[5] = t1[0];
    Y[1] = t1[1];
    Y[3] = mk<R>(t1[2].x, -t1[2].y);
    Y[7] = mk<R>(t1[3].x, -t1[3].y);
    Y[9] = t1[4];
}

template <typename R>
SC_HD void rdft20_inv(const cx<R> (&Y)[11], R (&x)[20]) {
    R t0[5], t2[5];
    cx<R> t1[5];
    rradix5_inv(Y[0].x, mk<R>(Y[4].x, -Y[4].y), mk<R>(Y[8].x, -Y[8].y), t0[0], t0[1], t0[2], t0[3], t0[4]);
    rradix5_inv(Y[10].x, Y[6], Y[2], t2[0], t2[1], t2[2], t2[3], t2[4]);
    t1[0] = Y[5]; t1[1] = Y[1];
    t1[2] = mk<R>(Y[3].x, -Y[3].y);
    t1[3] = mk<R>(Y[7].x, -Y[7].y);
    t1[4] = Y[9];
    radix5<true>(t1[0], t1[1], t1[2], t1[3], t1[4]);
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) {
        // inverse 4-point over k1 with T3 = conj T1: a0 = T0+T2+2ReT1, a2 = T0+T2-2ReT1, a1 = T0-T2-2ImT1, a3 = T0-T2+2ImT1
        const R s = t0[n2] + t2[n2], d = t0[n2] - t2[n2];
        const R re2 = t1[n2].x + t1[n2].x, im2 = t1[n2].y + t1[n2].y;
        x[(0 + 4 * n2) % 20] = s + re2;
        x[(10 + 4 * n2) % 20] = s - re2;
        x[(5 + 4 * n2) % 20] = d - im2;
        x[(15 + 4 * n2) % 20] = d + im2;
    }
}

}  // namespace scdsp
