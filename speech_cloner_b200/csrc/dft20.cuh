// 20-point complex DFT held entirely in registers (Good-Thomas 4 x 5, no internal twiddles).
//
// Building block of the real FFT-400 used by both halves of the hot path:
//   STFT   of audio_lib.py:141-147 / :267   (librosa.stft  -> scipy FFT of n_fft = 400)
//   iSTFT  of audio_lib.py:260              (librosa.istft -> scipy inverse FFT)
// n_fft = 400 = 20 x 20 in every shipped hp/*.json (win_length_ms 25 @ 16 kHz).
#pragma once
#include <cuda_runtime.h>

namespace scdsp {

__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// a * b
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__host__ __device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}

// 4-point DFT, in place.  INV = false: kernel exp(-2*pi*i*n*k/4); INV = true: exp(+...).
template <bool INV>
__host__ __device__ __forceinline__ void radix4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 s0 = cadd(a0, a2), d0 = csub(a0, a2);
    const float2 s1 = cadd(a1, a3), d1 = csub(a1, a3);
    a0 = cadd(s0, s1);
    a2 = csub(s0, s1);
    // forward: X1 = d0 - i*d1, X3 = d0 + i*d1
    const float2 p = make_float2(d0.x + d1.y, d0.y - d1.x);
    const float2 m = make_float2(d0.x - d1.y, d0.y + d1.x);
    a1 = INV ? m : p;
    a3 = INV ? p : m;
}

// 5-point DFT, in place.
template <bool INV>
__host__ __device__ __forceinline__ void radix5(float2& a0, float2& a1, float2& a2, float2& a3, float2& a4) {
    constexpr float C1 = 0.30901699437494745f;   // cos(2*pi/5)
    constexpr float C2 = -0.8090169943749473f;   // cos(4*pi/5)
    constexpr float S1 = 0.9510565162951535f;    // sin(2*pi/5)
    constexpr float S2 = 0.5877852522924731f;    // sin(4*pi/5)
    const float2 t1 = cadd(a1, a4), t3 = csub(a1, a4);
    const float2 t2 = cadd(a2, a3), t4 = csub(a2, a3);
    const float2 m1 = make_float2(fmaf(C2, t2.x, fmaf(C1, t1.x, a0.x)), fmaf(C2, t2.y, fmaf(C1, t1.y, a0.y)));
    const float2 m2 = make_float2(fmaf(C1, t2.x, fmaf(C2, t1.x, a0.x)), fmaf(C1, t2.y, fmaf(C2, t1.y, a0.y)));
    const float2 q1 = make_float2(fmaf(S2, t4.x, S1 * t3.x), fmaf(S2, t4.y, S1 * t3.y));
    const float2 q2 = make_float2(fmaf(-S1, t4.x, S2 * t3.x), fmaf(-S1, t4.y, S2 * t3.y));
    a0 = make_float2(a0.x + t1.x + t2.x, a0.y + t1.y + t2.y);
    // forward: X1 = m1 - i*q1, X4 = m1 + i*q1, X2 = m2 - i*q2, X3 = m2 + i*q2
    const float2 x1 = make_float2(m1.x + q1.y, m1.y - q1.x);
    const float2 x4 = make_float2(m1.x - q1.y, m1.y + q1.x);
    const float2 x2 = make_float2(m2.x + q2.y, m2.y - q2.x);
    const float2 x3 = make_float2(m2.x - q2.y, m2.y + q2.x);
    a1 = INV ? x4 : x1;
    a4 = INV ? x1 : x4;
    a2 = INV ? x3 : x2;
    a3 = INV ? x2 : x3;
}

// Unnormalised 20-point DFT, natural order in and out, fully unrolled (register resident).
//   input  index n = (5*n1 + 4*n2) mod 20,  n1 in [0,4), n2 in [0,5)
//   output index k = (5*k1 + 16*k2) mod 20
template <bool INV>
__host__ __device__ __forceinline__ void dft20(float2 (&v)[20]) {
    float2 t[4][5];
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) {
        float2 a0 = v[(0 + 4 * n2) % 20], a1 = v[(5 + 4 * n2) % 20];
        float2 a2 = v[(10 + 4 * n2) % 20], a3 = v[(15 + 4 * n2) % 20];
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
