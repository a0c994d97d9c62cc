// CPU emulation of one FFT-400 "unit" (20 threads, one frame pair): runs the exact
// __host__ __device__ routines of csrc/fft400.cuh thread by thread and compares them with a
// naive float64 DFT.  Built and run by tests/test_host_fft400.py (no GPU needed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../speech_cloner_b200/csrc/fft400.cuh"

using namespace scdsp;

static void naive_rdft(const std::vector<double>& x, std::vector<double>& re, std::vector<double>& im) {
    re.assign(kBins, 0.0); im.assign(kBins, 0.0);
    for (int k = 0; k < kBins; ++k)
        for (int n = 0; n < kNfft; ++n) {
            const double a = -2.0 * M_PI * (double)((long)n * k % kNfft) / kNfft;
            re[k] += x[n] * cos(a); im[k] += x[n] * sin(a);
        }
}

template <typename R>
static int check(const char* name, double tol_pow, double tol_rt) {
    std::vector<cx<R>> w400(kNfft);
    for (int m = 0; m < kNfft; ++m) {
        const double a = -2.0 * M_PI * m / kNfft;
        w400[m] = mk<R>((R)cos(a), (R)sin(a));
    }
    srand(7);  // same data for both precisions
    std::vector<double> xa(kNfft), xb(kNfft);
    for (int n = 0; n < kNfft; ++n) {
        xa[n] = (rand() / (double)RAND_MAX - 0.5) * (0.5 - 0.5 * cos(2 * M_PI * n / kNfft));
        xb[n] = (rand() / (double)RAND_MAX - 0.5) * (0.5 - 0.5 * cos(2 * M_PI * n / kNfft)) + 0.3 * sin(0.31 * n);
    }
    std::vector<double> ar, ai, br, bi;
    naive_rdft(xa, ar, ai); naive_rdft(xb, br, bi);

    std::vector<cx<R>> slots(kUnitSlots);
    TwReg<R> tw[20];
    TwTab<R> twt[20];
    for (int j = 0; j < 20; ++j) { tw[j].load(w400.data(), j); twt[j].load(w400.data(), j); }

    // ---------------- forward + power
    for (int j = 0; j < 20; ++j) {
        cx<R> z[20];
        for (int n1 = 0; n1 < 20; ++n1) z[n1] = mk<R>((R)(0.5 * xa[20 * n1 + j]), (R)(0.5 * xb[20 * n1 + j]));
        if (j & 1) fwd_step1(z, tw[j], &slots[j]); else fwd_step1(z, twt[j], &slots[j]);   // both twiddle sources
    }
    std::vector<float> pa(kBins, -1.f), pb(kBins, -1.f);
    cx<R> V[20][20];
    for (int c = 0; c < 20; ++c) {
        fwd_step2(V[c], &slots[c * kSlotLd]);
        store_power(V[c], c, pa.data(), pb.data());
    }
    double pmax = 0, perr = 0;
    for (int k = 0; k < kBins; ++k) {
        const double ta = ar[k] * ar[k] + ai[k] * ai[k], tb = br[k] * br[k] + bi[k] * bi[k];
        pmax = fmax(pmax, fmax(ta, tb));
        perr = fmax(perr, fmax(fabs(pa[k] - ta), fabs(pb[k] - tb)));
    }
    printf("[%s] power max %.4g  max abs err %.3g  rel %.3g\n", name, pmax, perr, perr / pmax);
    int fail = perr / pmax > tol_pow;

    // ---------------- real-input step 1 / step 1' (two real 20-point DFTs) against the same references
    {
        std::vector<cx<R>> slots2(kUnitSlots);
        for (int j = 0; j < 20; ++j) {
            R a[20], b[20];
            for (int n1 = 0; n1 < 20; ++n1) { a[n1] = (R)xa[20 * n1 + j]; b[n1] = (R)xb[20 * n1 + j]; }
            fwd_step1_real(a, b, twt[j], &slots2[j]);
        }
        double serr = 0, smax = 0;
        for (int i = 0; i < kUnitSlots; ++i) {
            if (i % kSlotLd == 20) continue;                      // padding column
            serr = fmax(serr, fmax(fabs((double)slots2[i].x - (double)slots[i].x), fabs((double)slots2[i].y - (double)slots[i].y)));
            smax = fmax(smax, fmax(fabs((double)slots[i].x), fabs((double)slots[i].y)));
        }
        printf("[%s] real step 1 vs packed step 1: max slot diff %.3g (max %.3g)\n", name, serr, smax);
        fail |= serr / smax > tol_rt * 50 + 1e-15;
        std::vector<float> pa2(kBins, -1.f), pb2(kBins, -1.f);
        double rerr = 0;
        for (int c = 0; c < 20; ++c) {
            cx<R> v[20];
            fwd_step2(v, &slots2[c * kSlotLd]);
            store_power(v, c, pa2.data(), pb2.data());
            inv_step2(v, &slots2[c * kSlotLd]);
        }
        for (int k = 0; k < kBins; ++k) rerr = fmax(rerr, fmax(fabs(pa2[k] - pa[k]), fabs(pb2[k] - pb[k])));
        printf("[%s] real path power vs packed path power: max diff %.3g\n", name, rerr);
        fail |= rerr / pmax > tol_pow;
        double rt = 0;
        for (int j = 0; j < 20; ++j) {
            R a[20], b[20];
            inv_step1_real(a, b, tw[j], &slots2[j]);
            for (int n1 = 0; n1 < 20; ++n1) {
                rt = fmax(rt, fabs((double)a[n1] / 400.0 - xa[20 * n1 + j]));
                rt = fmax(rt, fabs((double)b[n1] / 400.0 - xb[20 * n1 + j]));
            }
        }
        printf("[%s] real path forward -> inverse round trip: max abs err %.3g\n", name, rt);
        fail |= rt > tol_rt;
    }

    // ---------------- forward -> gl_update with A = |X| (identity) -> inverse == input
    std::vector<float> ampa(kBins), ampb(kBins), pha(kBins), phb(kBins);
    for (int k = 0; k < kBins; ++k) {
        ampa[k] = (float)hypot(ar[k], ai[k]); ampb[k] = (float)hypot(br[k], bi[k]);
        pha[k] = (float)atan2(ai[k], ar[k]); phb[k] = (float)atan2(bi[k], br[k]);
    }
    for (int mode = 0; mode < 3; ++mode) {
        for (int c = 0; c < 20; ++c) {
            cx<R> u[20];
            if (mode == 2) {                       // plain forward -> inverse, any precision
                for (int i = 0; i < 20; ++i) u[i] = V[c][i];
            } else {
                cxf uf[20];
                if (mode == 0) {
                    for (int i = 0; i < 20; ++i) uf[i] = mk<float>((float)V[c][i].x, (float)V[c][i].y);
                    gl_update(uf, c, ampa.data(), ampb.data());
                } else {
                    gl_init_state(uf, c, ampa.data(), ampb.data(), pha.data(), phb.data());
                }
                for (int i = 0; i < 20; ++i) u[i] = mk<R>((R)uf[i].x, (R)uf[i].y);
            }
            inv_step2(u, &slots[c * kSlotLd]);
        }
        double rerr = 0, xmax = 0;
        for (int j = 0; j < 20; ++j) {
            cx<R> h[20];
            if (j & 1) inv_step1(h, twt[j], &slots[j]); else inv_step1(h, tw[j], &slots[j]);
            for (int n1 = 0; n1 < 20; ++n1) {
                rerr = fmax(rerr, fabs(h[n1].x / 400.0 - xa[20 * n1 + j]));
                rerr = fmax(rerr, fabs(h[n1].y / 400.0 - xb[20 * n1 + j]));
                xmax = fmax(xmax, fmax(fabs(xa[20 * n1 + j]), fabs(xb[20 * n1 + j])));
            }
        }
        printf("[%s] mode %d roundtrip max abs err %.3g (signal max %.3g)\n", name, mode, rerr, xmax);
        fail |= rerr / xmax > (mode == 2 ? tol_rt : 5e-6);
    }
    return fail;
}

int main() {
    int fail = check<float>("float32", 2e-6, 5e-6);
    fail |= check<double>("float64", 1e-7, 1e-14);   // power is stored as float32
    printf(fail ? "FAIL\n" : "OK\n");
    return fail;
}
