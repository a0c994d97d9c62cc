// C ABI (include/speechdsp.h) + plan / host-side tables for the speech-cloner DSP hot path.
// Single translation unit: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -shared.
#include "../../include/speechdsp.h"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fe_kernels.cuh"
#include "fe_ws.cuh"
#include "fe_split.cuh"
#include "gl_kernels.cuh"
#include "generic_kernels.cuh"
#include "phn_kernels.cuh"

using namespace scdsp;

// ------------------------------------------------------------------------------------- errors
static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define SC_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? SC_ERR_NO_DEVICE : SC_ERR_CUDA, \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                   \
    } while (0)
#define SC_LAUNCHED()                                                                          \
    do {                                                                                       \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                    \
        SC_CUDA(cudaGetLastError());                                                           \
    } while (0)

// ------------------------------------------------------------------------------ device buffers
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return fail(SC_ERR_CUDA, std::string("cudaMalloc workspace: ") + cudaGetErrorString(e));
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// host blob that is copied to the device in one cudaMemcpyAsync; offsets are 16-byte aligned
struct Blob {
    std::vector<unsigned char> bytes;
    template <typename T> size_t add(const T* src, size_t n) {
        size_t off = (bytes.size() + 15) & ~size_t(15);
        bytes.resize(off + n * sizeof(T));
        if (n) memcpy(bytes.data() + off, src, n * sizeof(T));
        return off;
    }
    template <typename T> size_t add(const std::vector<T>& v) { return add(v.data(), v.size()); }
};

// descriptor staging: pinned host blob -> one async copy; the event guards reuse of the blob
struct DescStage {
    DevBuf desc;
    unsigned char* pinned = nullptr;
    size_t pinned_cap = 0;
    cudaEvent_t done = nullptr;
    bool pending = false;
    void release() {
        if (pending && done) cudaEventSynchronize(done);
        if (done) cudaEventDestroy(done);
        if (pinned) cudaFreeHost(pinned);
        desc.release();
        pinned = nullptr; pinned_cap = 0; done = nullptr; pending = false;
    }
};

struct sc_plan {
    sc_params prm{};
    int n_bins = 0;
    bool fast = false;
    int device = 0;
    // constant tables (one allocation)
    DevBuf tables;
    const cxf* w400 = nullptr;
    const cxd* w400_d = nullptr;
    const double* fe_win_half_d = nullptr;
    const float* fe_win_half = nullptr;
    const float* gl_win_half = nullptr;
    const float* gl_win_inv = nullptr;
    const double* gl_win_sq = nullptr;
    const float* gl_inv_wss = nullptr;
    const float* zero_row = nullptr;
    const float2* mel_w = nullptr;
    const int32_t* mel_istart = nullptr;
    const int32_t* mel_chunk = nullptr;
    const float* dct_e = nullptr;
    const float* dct_o = nullptr;
    // generic-size tables
    const double* g_fe_win = nullptr;    // analysis window (n_fft), float64
    const float* g_gl_win = nullptr;     // hann (n_fft)
    const cxf* g_wn = nullptr;           // exp(-2*pi*i*m/n_fft)
    const cxd* g_wn_d = nullptr;
    const double* g_win_sq = nullptr;
    bool fp32_fft = false;               // sc_params.fft_precision == 1
    // host copies used by generic paths / tests
    std::vector<double> fe_window, gl_window;
    // per-call workspaces
    DescStage ds;
    DevBuf work, work2;
    // optional per-kernel timing (bench.py's roofline): events recorded on the caller's stream
    bool profile = false;
    std::vector<cudaEvent_t> pev;   // 4 events per front-end group, or 3 for a Griffin-Lim call
    int pev_groups = 0;
    int pev_kind = 0;        // 1 = front-end (gain | pass A | pass B), 2 = Griffin-Lim (init | iterations)
    int pev_iters = 0;
    WsMelParam ws_mel{};     // band / bin ranges of the epilogue warps of k_fe_pass_a_ws
    const int4* ws_brec = nullptr;  // per-band (first tap, float4 blocks, weight offset)
    const float* ws_wt = nullptr;   // padded filterbank weights
    bool use_b2 = true;      // vector form of pass B (env SC_FE_B2=0 selects the scalar kernel)
    bool use_b3 = false;     // compile-time specialised pass B (80 mels, 40 MFCC, 201 bins; env SC_FE_B3=0 disables)
    bool use_split = false;  // pass A as FFT-only kernel + streaming mel kernel (env SC_FE_SPLIT=1)
    bool use_ws = true;      // warp-specialised pass A (env SC_FE_WS=0 selects the older persistent kernel)
    int64_t fe_group_frames = int64_t(1) << 60;   // frames per front-end group (env SC_FE_GROUP_FRAMES); measured: grouping for L2 residency only adds launch latency, so off by default
};

static int upload_blob(DescStage& ds, const Blob& b, cudaStream_t st) {
    if (!ds.done) SC_CUDA(cudaEventCreateWithFlags(&ds.done, cudaEventDisableTiming));
    if (ds.pending) {
        SC_CUDA(cudaEventSynchronize(ds.done));
        ds.pending = false;
    }
    if (b.bytes.size() > ds.pinned_cap) {
        if (ds.pinned) cudaFreeHost(ds.pinned);
        ds.pinned = nullptr; ds.pinned_cap = 0;
        size_t want = b.bytes.size() * 2 + 4096;
        SC_CUDA(cudaMallocHost((void**)&ds.pinned, want));
        ds.pinned_cap = want;
    }
    if (int rc = ds.desc.ensure(b.bytes.size())) return rc;
    memcpy(ds.pinned, b.bytes.data(), b.bytes.size());
    SC_CUDA(cudaMemcpyAsync(ds.desc.p, ds.pinned, b.bytes.size(), cudaMemcpyHostToDevice, st));
    SC_CUDA(cudaEventRecord(ds.done, st));
    ds.pending = true;
    return 0;
}
static int upload_blob(sc_plan* pl, const Blob& b, cudaStream_t st) { return upload_blob(pl->ds, b, st); }
template <typename T> static const T* at(const DescStage& ds, size_t off) {
    return reinterpret_cast<const T*>(static_cast<const unsigned char*>(ds.desc.p) + off);
}
template <typename T> static const T* at(const sc_plan* pl, size_t off) { return at<T>(pl->ds, off); }

// ------------------------------------------------------------------------------ host tables
static std::vector<double> hann_periodic(int n) {
    std::vector<double> w(n);
    for (int i = 0; i < n; ++i) w[i] = 0.5 - 0.5 * cos(2.0 * M_PI * i / n);
    return w;
}
static std::vector<double> pad_center(const std::vector<double>& w, int n_fft) {
    std::vector<double> out(n_fft, 0.0);
    const int lpad = (n_fft - (int)w.size()) / 2;
    for (size_t i = 0; i < w.size(); ++i) out[lpad + i] = w[i];
    return out;
}
// librosa 0.6 hz_to_mel / mel_to_hz, htk=False (Slaney)
static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

struct MelSparse {
    std::vector<float2> w;          // per bin (up, dn)
    std::vector<int32_t> istart;    // n_mels + 2
    std::vector<int32_t> chunk;     // kMaxMelChunks + 1
};

// librosa.filters.mel(sr, n_fft, n_mels, fmin=0, fmax=sr/2, htk=False, norm=1) (audio_lib.py:160-166)
// in the per-bin form: bin k lies in mel interval i(k) = [edge_i, edge_{i+1}) and feeds only the
// rising slope of band i and the falling slope of band i-1.
static MelSparse build_mel(int sr, int n_fft, int n_mels) {
    const int n_bins = 1 + n_fft / 2;
    const double fmax = sr / 2.0;
    std::vector<double> edge(n_mels + 2);
    const double m_lo = hz_to_mel(0.0), m_hi = hz_to_mel(fmax);
    const double step = (m_hi - m_lo) / (n_mels + 1);
    for (int i = 0; i < n_mels + 2; ++i) edge[i] = mel_to_hz(i == n_mels + 1 ? m_hi : m_lo + i * step);
    auto weight = [&](int band, double f) -> double {
        if (band < 0 || band >= n_mels) return 0.0;
        const double lower = (f - edge[band]) / (edge[band + 1] - edge[band]);
        const double upper = (edge[band + 2] - f) / (edge[band + 2] - edge[band + 1]);
        const double t = fmax > 0 ? std::fmax(0.0, std::fmin(lower, upper)) : 0.0;
        return t * (2.0 / (edge[band + 2] - edge[band]));
    };
    MelSparse ms;
    ms.w.resize(n_bins);
    ms.istart.assign(n_mels + 2, n_bins);
    int prev_i = -1;
    for (int k = 0; k < n_bins; ++k) {
        const double f = (n_bins > 1) ? k * fmax / (n_bins - 1) : 0.0;
        int i = 0;
        while (i < n_mels && edge[i + 1] <= f) ++i;     // edge[i] <= f < edge[i+1], clamped to n_mels
        ms.w[k] = make_float2((float)weight(i, f), (float)weight(i - 1, f));
        for (int q = prev_i + 1; q <= i; ++q) ms.istart[q] = k;
        prev_i = i;
    }
    for (int q = prev_i + 1; q <= n_mels + 1; ++q) ms.istart[q] = n_bins;
    // balance the band chunks by cost
    std::vector<double> cost(n_mels + 1);
    double total = 0;
    double band_cost = 20.0;                                  // instructions per closed band relative to 3 per bin
    if (const char* e = getenv("SC_MEL_BAND_COST")) band_cost = atof(e);
    for (int i = 0; i <= n_mels; ++i) { cost[i] = 3.0 * (ms.istart[i + 1] - ms.istart[i]) + band_cost; total += cost[i]; }
    ms.chunk.assign(kMaxMelChunks + 1, n_mels);
    ms.chunk[0] = 0;
    double run = 0; int c = 1;
    for (int i = 0; i < n_mels && c < kMaxMelChunks; ++i) {
        run += cost[i];
        if (run >= total * c / kMaxMelChunks) ms.chunk[c++] = i + 1;
    }
    for (; c <= kMaxMelChunks; ++c) ms.chunk[c] = n_mels;
    return ms;
}

static std::vector<float> build_inv_wss(const std::vector<double>& win, int n_fft, int hop) {
    std::vector<float> out(hop);
    for (int r = 0; r < hop; ++r) {
        float acc = 0.f;
        // covering frames in ascending order = window index descending
        int top = r + ((n_fft - 1 - r) / hop) * hop;
        for (int idx = top; idx >= 0; idx -= hop) acc = (float)((double)acc + win[idx] * win[idx]);
        out[r] = acc > 1.1754944e-38f ? 1.0f / acc : 1.0f;
    }
    return out;
}

// ------------------------------------------------------------------------------------- plan
extern "C" int sc_plan_create(const sc_params* p, sc_plan** out) {
    if (!p || !out) return fail(SC_ERR_INVALID, "sc_plan_create: null argument");
    *out = nullptr;
    if (p->n_fft < 2 || (p->n_fft & 1)) return fail(SC_ERR_INVALID, "n_fft must be even and >= 2");
    if (p->n_fft > kGenMaxNfft) return fail(SC_ERR_UNSUPPORTED, "n_fft larger than 2048 is not supported");
    if (p->win_length < 1 || p->win_length > p->n_fft) return fail(SC_ERR_INVALID, "win_length must be in [1, n_fft]");
    if (p->hop_length < 1) return fail(SC_ERR_INVALID, "hop_length must be >= 1");
    if (p->n_mels < 1 || p->n_mels > kMaxMels) return fail(SC_ERR_INVALID, "n_mels must be in [1, 128]");
    if (p->n_mfcc < 1 || p->n_mfcc > p->n_mels) return fail(SC_ERR_INVALID, "n_mfcc must be in [1, n_mels]");
    if (p->sample_rate < 1) return fail(SC_ERR_INVALID, "sample_rate must be positive");
    if (p->fft_precision != 0 && p->fft_precision != 1) return fail(SC_ERR_INVALID, "fft_precision must be 0 (float64) or 1 (float32)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SC_ERR_NO_DEVICE, "no CUDA device: speechdsp has no CPU fallback");

    sc_plan* pl = new sc_plan();
    pl->prm = *p;
    pl->prm.window_host = nullptr;
    pl->n_bins = 1 + p->n_fft / 2;
    pl->fast = (p->n_fft == kNfft && p->hop_length == kHop);
    pl->fp32_fft = p->fft_precision == 1;
    if (const char* e = getenv("SC_FE_GROUP_FRAMES")) { const long long v = atoll(e); if (v > 0) pl->fe_group_frames = v; }
    cudaGetDevice(&pl->device);

    std::vector<double> w = p->window_host ? std::vector<double>(p->window_host, p->window_host + p->win_length)
                                           : hann_periodic(p->win_length);
    pl->fe_window = pad_center(w, p->n_fft);
    pl->gl_window = pad_center(hann_periodic(p->win_length), p->n_fft);   // librosa default inside GL (:260, :267)

    const int n_fft = p->n_fft;
    Blob b;
    std::vector<cxf> wn(n_fft);
    std::vector<cxd> wn_d(n_fft);
    for (int m = 0; m < n_fft; ++m) {
        const double a = -2.0 * M_PI * m / n_fft;
        wn_d[m] = mk<double>(cos(a), sin(a));
        wn[m] = mk<float>((float)wn_d[m].x, (float)wn_d[m].y);
    }
    std::vector<float> fe_half(n_fft), gl_half(n_fft), gl_inv(n_fft), gl_w(n_fft);
    std::vector<double> gl_sq(n_fft), fe_half_d(n_fft);
    for (int i = 0; i < n_fft; ++i) {
        fe_half[i] = (float)(0.5 * pl->fe_window[i]);
        gl_half[i] = (float)(0.5 * pl->gl_window[i]);
        gl_inv[i] = (float)(pl->gl_window[i] / n_fft);
        gl_sq[i] = pl->gl_window[i] * pl->gl_window[i];
        fe_half_d[i] = 0.5 * pl->fe_window[i];
        gl_w[i] = (float)pl->gl_window[i];
    }
    std::vector<float> inv_wss = build_inv_wss(pl->gl_window, n_fft, p->hop_length);
    MelSparse ms = build_mel(p->sample_rate, n_fft, p->n_mels);
    std::vector<int4> ws_brec(kWsMaxPairs + 1, make_int4(0, 0, 0, 0));
    std::vector<float> ws_wt;
    if (pl->n_bins == kBins) {
        // filterbank of k_fe_pass_a_ws (see WsMelParam): band b = rising slope over interval b (ms.w[k].x) +
        // falling slope over interval b+1 (ms.w[k].y); warps get contiguous band ranges balanced by cost and walk
        // them two bands at a time, both zero padded to the same number of float4 blocks
        const int nm = p->n_mels;
        auto blocks_of = [&](int bnd) { return (ms.istart[bnd + 2] - ms.istart[bnd] + 3) / 4; };
        auto weight_of = [&](int bnd, int q) -> float {
            const int k = ms.istart[bnd] + q;
            return k < ms.istart[bnd + 1] ? ms.w[k].x : (k < ms.istart[bnd + 2] ? ms.w[k].y : 0.f);
        };
        std::vector<double> cost(nm);
        double total = 0;
        // measured per band pair: ~350 cycles + ~105 per float4 block (role timing of the SC_WS_DEBUG build)
        double ws_cost_block = 1.0, ws_cost_band = 3.4;
        if (const char* e = getenv("SC_WS_COST_BAND")) ws_cost_band = atof(e);
        for (int bnd = 0; bnd < nm; ++bnd) { cost[bnd] = ws_cost_block * blocks_of(bnd) + ws_cost_band; total += cost[bnd]; }
        WsMelParam& wm = pl->ws_mel;
        wm.n_mels = nm;
        wm.chunk[0] = 0;
        double run = 0; int c = 1;
        for (int i = 0; i < nm && c < kWsEpiWarps; ++i) {
            run += cost[i];
            if (run >= total * c / kWsEpiWarps) wm.chunk[c++] = i + 1;
        }
        for (; c <= kWsEpiWarps; ++c) wm.chunk[c] = nm;
        int n_pairs = 0;
        for (int w = 0; w < kWsEpiWarps; ++w) {
            wm.pair0[w] = n_pairs;
            for (int bnd = wm.chunk[w]; bnd < wm.chunk[w + 1]; bnd += 2) {
                const bool has_b = bnd + 1 < wm.chunk[w + 1];
                const int nb = std::max(blocks_of(bnd), has_b ? blocks_of(bnd + 1) : 0);
                ws_brec[n_pairs++] = make_int4(ms.istart[bnd], has_b ? ms.istart[bnd + 1] : ms.istart[bnd], nb, (int)ws_wt.size() / 4);
                for (int blk = 0; blk < nb; ++blk) {
                    for (int q = 0; q < 4; ++q) ws_wt.push_back(weight_of(bnd, 4 * blk + q));
                    for (int q = 0; q < 4; ++q) ws_wt.push_back(has_b ? weight_of(bnd + 1, 4 * blk + q) : 0.f);
                }
            }
        }
        wm.pair0[kWsEpiWarps] = n_pairs;
        wm.n_taps = (int)ws_wt.size();
        if (wm.n_taps > kWsMaxTaps || n_pairs > kWsMaxPairs) pl->use_ws = false;   // not reachable for 201 bins and <= 128 bands
    }
    if (ws_wt.empty()) ws_wt.push_back(0.f);
    if (const char* e = getenv("SC_FE_WS")) pl->use_ws = pl->use_ws && atoi(e) != 0;
    if (const char* e = getenv("SC_FE_B2")) pl->use_b2 = atoi(e) != 0;
    if (const char* e = getenv("SC_FE_SPLIT")) pl->use_split = atoi(e) != 0;
    // librosa.filters.dct (audio_lib.py:176): row 0 = 1/sqrt(N), row q = sqrt(2/N) cos(q (2n+1) pi / 2N),
    // split into even / odd rows over the first half of the inputs (see k_fe_pass_b)
    const FbLayout fbl = fb_layout(p->n_mels, p->n_mfcc);
    std::vector<float> dct_e((size_t)fbl.half * fbl.ne_pad, 0.f), dct_o((size_t)fbl.half * fbl.no_pad, 0.f);
    for (int q = 0; q < p->n_mfcc; ++q)
        for (int m = 0; m < fbl.half; ++m) {
            const double v = q == 0 ? 1.0 / sqrt((double)p->n_mels)
                                    : cos(q * (2.0 * m + 1.0) * M_PI / (2.0 * p->n_mels)) * sqrt(2.0 / p->n_mels);
            if (q & 1) dct_o[(size_t)m * fbl.no_pad + q / 2] = (float)v;
            else dct_e[(size_t)m * fbl.ne_pad + q / 2] = (float)v;
        }
    if (pl->n_bins == kBins && p->n_mels == kB3Mels && p->n_mfcc == kB3Mfcc) {
        pl->use_b3 = true;
        if (const char* e = getenv("SC_FE_B3")) pl->use_b3 = atoi(e) != 0;
    }
    const size_t o_wn = b.add(wn), o_feh = b.add(fe_half), o_glh = b.add(gl_half), o_gli = b.add(gl_inv);
    const size_t o_sq = b.add(gl_sq), o_wss = b.add(inv_wss), o_mw = b.add(ms.w), o_mi = b.add(ms.istart);
    const size_t o_mc = b.add(ms.chunk), o_dcte = b.add(dct_e), o_dcto = b.add(dct_o), o_few = b.add(pl->fe_window), o_glw = b.add(gl_w);
    const size_t o_wnd = b.add(wn_d), o_fehd = b.add(fe_half_d), o_wsrec = b.add(ws_brec), o_wswt = b.add(ws_wt);
    std::vector<float> zeros(pl->n_bins, 0.f);
    const size_t o_zero = b.add(zeros);
    if (pl->tables.ensure(b.bytes.size())) { delete pl; return SC_ERR_CUDA; }
    e = cudaMemcpy(pl->tables.p, b.bytes.data(), b.bytes.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        pl->tables.release();
        delete pl;
        return fail(SC_ERR_CUDA, std::string("plan upload: ") + cudaGetErrorString(e));
    }
    const unsigned char* base = static_cast<const unsigned char*>(pl->tables.p);
    pl->w400 = (const cxf*)(base + o_wn);          pl->g_wn = pl->w400;
    pl->w400_d = (const cxd*)(base + o_wnd);       pl->g_wn_d = pl->w400_d;
    pl->fe_win_half_d = (const double*)(base + o_fehd);
    pl->fe_win_half = (const float*)(base + o_feh);
    pl->gl_win_half = (const float*)(base + o_glh);
    pl->gl_win_inv = (const float*)(base + o_gli);
    pl->gl_win_sq = (const double*)(base + o_sq);  pl->g_win_sq = pl->gl_win_sq;
    pl->gl_inv_wss = (const float*)(base + o_wss);
    pl->zero_row = (const float*)(base + o_zero);
    pl->mel_w = (const float2*)(base + o_mw);
    pl->mel_istart = (const int32_t*)(base + o_mi);
    pl->mel_chunk = (const int32_t*)(base + o_mc);
    pl->dct_e = (const float*)(base + o_dcte);
    pl->dct_o = (const float*)(base + o_dcto);
    pl->ws_brec = (const int4*)(base + o_wsrec);
    pl->ws_wt = (const float*)(base + o_wswt);
    pl->g_fe_win = (const double*)(base + o_few);
    pl->g_gl_win = (const float*)(base + o_glw);
    *out = pl;
    return SC_OK;
}

extern "C" void sc_plan_destroy(sc_plan* pl) {
    if (!pl) return;
    pl->ds.release();
    for (auto& e : pl->pev) if (e) cudaEventDestroy(e);
    pl->tables.release(); pl->work.release(); pl->work2.release();
    delete pl;
}

extern "C" int sc_plan_is_fast_path(const sc_plan* pl) { return pl && pl->fast ? 1 : 0; }

extern "C" int64_t sc_num_frames(const sc_plan* pl, int64_t n) { return pl ? 1 + n / pl->prm.hop_length : 0; }

static int prof_event(sc_plan* pl, int idx, cudaStream_t st) {
    while ((int)pl->pev.size() <= idx) {
        cudaEvent_t e = nullptr;
        SC_CUDA(cudaEventCreate(&e));
        pl->pev.push_back(e);
    }
    SC_CUDA(cudaEventRecord(pl->pev[idx], st));
    return 0;
}

static FeTables fe_tables(const sc_plan* pl) {
    FeTables t;
    t.w400 = pl->w400; t.win_half = pl->fe_win_half; t.w400_d = pl->w400_d; t.win_half_d = pl->fe_win_half_d;
    t.mel_w = pl->mel_w; t.mel_istart = pl->mel_istart;
    t.mel_chunk = pl->mel_chunk; t.dct_e = pl->dct_e; t.dct_o = pl->dct_o;
    t.n_mels = pl->prm.n_mels; t.n_mfcc = pl->prm.n_mfcc;
    return t;
}
static FeParams fe_params(const sc_plan* pl) {
    const sc_params& p = pl->prm;
    FeParams f;
    f.pre_emphasis = p.pre_emphasis;
    f.mean_abs_amp_norm = p.mean_abs_amp_norm;
    f.mfcc_norm_factor = (float)p.mfcc_norm_factor;
    f.m_db_norm_factor = (float)p.m_db_norm_factor;
    f.p_db_norm_factor = (float)p.p_db_norm_factor;
    f.use_gain = p.mean_abs_amp_norm != 1.0;
    f.norm_first = p.mfcc_normalize_first != 0;
    f.use_delta = p.calc_mfcc_derivative != 0;
    f.clip = p.clip_output != 0;
    f.shift_p = p.p_db_norm_factor != 1.0;
    f.shift_m = p.m_db_norm_factor != 1.0;
    return f;
}

// ------------------------------------------------------------------------------ front-end
// (utterance, subtree index) of every CTA of k_abs_pairwise2
static std::vector<int2> abs_recs(const std::vector<int32_t>& pre, int n) {
    std::vector<int2> r((size_t)pre[n]);
    for (int u = 0; u < n; ++u)
        for (int i = pre[u]; i < pre[u + 1]; ++i) r[i] = make_int2(u, i - pre[u]);
    return r;
}
// leaf pass of the |y| sum: 4 = persistent staged (default on the fast path), 3 = staged, thread per leaf;
// 2 = 8 lanes per slot; 1 = shared-memory heap
static int abs_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SC_FE_ABS"); v = e ? atoi(e) : 4; }
    return v;
}
// one L2-resident group of utterances [0, n) (pointers already offset by the caller)
static int frontend_range(sc_plan* pl, const float* wav, const int64_t* soff, const int64_t* slen_in, int32_t n,
                          float* mfcc, float* mel, float* pdb, const int64_t* foff, cudaStream_t st, int group) {
    const int hop = pl->prm.hop_length;
    std::vector<int64_t> slen(slen_in, slen_in + n), so(soff, soff + n), fo(foff, foff + n);
    std::vector<int32_t> fcnt(n), pre_abs(n + 1), pre_a(n + 1), pre_b(n + 1), pre_b3(n + 1), pre_int(n + 1), pre_mel(n + 1), ifirst(n), icount(n);
    std::vector<int64_t> heap_off(n + 1);
    int64_t total_frames_span = 0;
    constexpr int kPU = 8;                         // units per CTA of the fast pass-A kernels (16 frames per tile)
    const bool ws = pl->fast && pl->use_ws;        // long utterances: warp-specialised kernel, 24-frame tiles
    const bool split = ws && pl->use_split && SC_DB_IN_PASS_B;   // ... or the FFT-only + mel kernel pair, 12-frame tiles
    const int ws_frames = split ? kSpFrames : kWsFrames;
    const int a_frames = pl->fast ? 2 * kPU : kGenFeFrames;
    for (int u = 0; u < n; ++u) {
        const int64_t T = 1 + slen[u] / hop;
        fcnt[u] = (int32_t)T;
        if (fo[u] + T > total_frames_span) total_frames_span = fo[u] + T;
    }
    pre_abs[0] = pre_a[0] = pre_b[0] = pre_b3[0] = pre_int[0] = pre_mel[0] = 0;
    heap_off[0] = 0;
    for (int u = 0; u < n; ++u) {
        const int64_t ta = pre_abs[u] + (int64_t(1) << abs_depth(slen[u]));
        heap_off[u + 1] = heap_off[u] + (int64_t(2) << abs_depth(slen[u]));
        // fast path: every tile of an utterance at least two tile spans long runs in the persistent kernel (its
        // first / last tiles gather their reflect padding); shorter utterances, whose padding could wrap more
        // than once, use the plain kernel
        int64_t n_tiles_u = (fcnt[u] + a_frames - 1) / a_frames, k_lo = 0, k_hi = -1;
        int64_t n_int = 0, n_mel = 0;
        if (pl->fast) {
            const int p_frames = ws ? ws_frames : a_frames;                   // tile of the persistent kernel
            const int64_t span = (int64_t)kHop * (p_frames - 1) + kNfft;
            if (slen[u] >= 2 * span + 16) {
                k_hi = n_tiles_u - 1;
                n_int = ws ? (fcnt[u] + ws_frames - 1) / ws_frames : n_tiles_u;
                n_mel = (fcnt[u] + kMelFrames - 1) / kMelFrames;
            }
        }
        ifirst[u] = (int32_t)k_lo; icount[u] = (int32_t)(k_hi >= k_lo ? k_hi - k_lo + 1 : 0);
        pre_int[u + 1] = (int32_t)(pre_int[u] + n_int);
        pre_mel[u + 1] = (int32_t)(pre_mel[u] + n_mel);
        const int64_t tb = pre_a[u] + n_tiles_u - icount[u];                  // edge (or all generic-path) tiles
        const int64_t tc = pre_b[u] + (fcnt[u] + kFbFrames - 1) / kFbFrames;
        if (ta > INT32_MAX || tb > INT32_MAX) return fail(SC_ERR_INVALID, "sc_frontend_batch: batch too large");
        pre_abs[u + 1] = (int32_t)ta; pre_a[u + 1] = (int32_t)tb; pre_b[u + 1] = (int32_t)tc;
        pre_b3[u + 1] = pre_b3[u] + (fcnt[u] + kB3Frames - 1) / kB3Frames;
    }
    Blob b;
    const size_t o_so = b.add(so), o_sl = b.add(slen), o_fo = b.add(fo), o_fc = b.add(fcnt);
    const size_t o_pabs = b.add(pre_abs), o_pa = b.add(pre_a), o_pb = b.add(pre_b), o_heap = b.add(heap_off);
    const size_t o_pint = b.add(pre_int), o_if = b.add(ifirst), o_ic = b.add(icount);
    const size_t o_arec = b.add(abs_recs(pre_abs, n));
    const size_t o_pb3 = b.add(pre_b3), o_pmel = b.add(pre_mel);
    if (int rc = upload_blob(pl, b, st)) return rc;

    // workspace: stats | abs partials | raw mel
    const size_t w_stat = 0;
    const size_t w_part = (sizeof(UttStat) * n + 255) & ~size_t(255);
    const size_t w_mel = (w_part + sizeof(float) * (size_t)heap_off[n] + 255) & ~size_t(255);
    const size_t w_tiles = (w_mel + sizeof(float) * (size_t)total_frames_span * pl->prm.n_mels + 255) & ~size_t(255);
    const size_t w_b3 = (w_tiles + (ws ? sizeof(WsTile) * (size_t)pre_int[n] : 0) + 255) & ~size_t(255);
    const size_t w_arec = (w_b3 + sizeof(B3Tile) * (size_t)pre_b3[n] + 255) & ~size_t(255);
    const size_t w_end = w_arec + sizeof(AbsRec) * (size_t)pre_abs[n];
    if (int rc = pl->work.ensure(w_end)) return rc;
    unsigned char* wb = static_cast<unsigned char*>(pl->work.p);
    UttStat* stat = reinterpret_cast<UttStat*>(wb + w_stat);
    float* heap = reinterpret_cast<float*>(wb + w_part);
    float* mel_raw = reinterpret_cast<float*>(wb + w_mel);

    Ragged rg;
    rg.sample_off = at<int64_t>(pl, o_so); rg.sample_len = at<int64_t>(pl, o_sl);
    rg.frame_off = at<int64_t>(pl, o_fo); rg.frame_cnt = at<int32_t>(pl, o_fc);
    rg.n_utts = n;
    rg.int_first = nullptr; rg.int_count = nullptr;
    const FeTables tb = fe_tables(pl);
    const FeParams fp = fe_params(pl);

    if (pl->profile) { if (int rc = prof_event(pl, 4 * group, st)) return rc; pl->pev_kind = 1; pl->pev_groups = group + 1; }
    // one launch writes every per-call table of the fast path (sub-tree records, pass A tiles, pass B tiles)
    const bool b3 = pl->fast && pl->use_b3 && pl->use_b2;
    const bool setup = pl->fast && (ws || b3 || abs_variant() == 4);
    static int n_sm_all = 0;
    if (!n_sm_all) SC_CUDA(cudaDeviceGetAttribute(&n_sm_all, cudaDevAttrMultiProcessorCount, pl->device));
    if (setup) {
        SetupArgs sa;
        sa.pre_abs = at<int32_t>(pl, o_pabs); sa.pre_ws = at<int32_t>(pl, o_pint); sa.pre_b3 = at<int32_t>(pl, o_pb3);
        sa.heap_off = at<int64_t>(pl, o_heap);
        sa.n_abs = (fp.use_gain && abs_variant() == 4) ? pre_abs[n] : 0;
        sa.n_ws = ws ? pre_int[n] : 0;
        sa.ws_frames = ws_frames;
        sa.n_b3 = b3 ? pre_b3[n] : 0;
        sa.abs_out = reinterpret_cast<AbsRec*>(wb + w_arec);
        sa.ws_out = reinterpret_cast<WsTile*>(wb + w_tiles);
        sa.b3_out = reinterpret_cast<B3Tile*>(wb + w_b3);
        const int total = sa.n_abs + sa.n_ws + sa.n_b3;
        if (total > 0) {
            k_fe_setup<<<(total + 255) / 256, 256, 0, st>>>(rg, sa);
            SC_LAUNCHED();
        }
    }
    if (fp.use_gain) {
        rg.tile_prefix = at<int32_t>(pl, o_pabs);
        if (pl->fast && abs_variant() == 4) {
            const int grid = pre_abs[n] < 4 * n_sm_all ? pre_abs[n] : 4 * n_sm_all;
            k_abs_pairwise4<<<grid, kAbs3Threads, 0, st>>>(wav, reinterpret_cast<const AbsRec*>(wb + w_arec), pre_abs[n], heap);
        } else if (abs_variant() >= 3) k_abs_pairwise3<<<pre_abs[n], kAbs3Threads, 0, st>>>(wav, rg, at<int2>(pl, o_arec), at<int64_t>(pl, o_heap), heap);
        else if (abs_variant() == 2) k_abs_pairwise2<<<pre_abs[n], kAbs2Threads, 0, st>>>(wav, rg, at<int2>(pl, o_arec), at<int64_t>(pl, o_heap), heap);
        else k_abs_pairwise<<<pre_abs[n], kAbsThreads, 0, st>>>(wav, rg, at<int64_t>(pl, o_heap), heap);
        SC_LAUNCHED();
    }
    rg.tile_prefix = at<int32_t>(pl, o_pabs);
    k_gain_finalize<<<(n + 3) / 4, 128, 0, st>>>(rg, at<int64_t>(pl, o_heap), heap, stat, fp.mean_abs_amp_norm, fp.use_gain, nullptr);
    SC_LAUNCHED();

    if (pl->profile) if (int rc = prof_event(pl, 4 * group + 1, st)) return rc;
    rg.tile_prefix = at<int32_t>(pl, o_pa);
    if (pl->fast) {
        static bool attr_set = false;
        static int n_sm = 0;
        if (!attr_set) {
            SC_CUDA(cudaFuncSetAttribute(k_fe_pass_a<float, kPU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            SC_CUDA(cudaFuncSetAttribute(k_fe_pass_a<double, kPU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
            SC_CUDA(cudaFuncSetAttribute(k_fe_pass_a_persist<float, kPU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            SC_CUDA(cudaFuncSetAttribute(k_fe_pass_a_persist<double, kPU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
            SC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, pl->device));
            attr_set = true;
        }
        const size_t mel_bytes = sizeof(float) * 2 * kPU * (pl->prm.n_mels + 1);
        rg.int_first = at<int32_t>(pl, o_if); rg.int_count = at<int32_t>(pl, o_ic);
        if (split && pre_int[n] > 0) {
            // FFT-only persistent kernel (three prep + FFT pipelines per SM) followed by the streaming mel kernel
            static bool sp_attr = false;
            if (!sp_attr) {
                SC_CUDA(cudaFuncSetAttribute(k_fe_fft<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpSmem<float>)));
                SC_CUDA(cudaFuncSetAttribute(k_fe_fft<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpSmem<double>)));
                SC_CUDA(cudaFuncSetAttribute(k_fe_mel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fe_mel_smem_bytes(kMaxMels)));
                sp_attr = true;
            }
            WsTile* tiles = reinterpret_cast<WsTile*>(wb + w_tiles);
            const int need = (pre_int[n] + kSpPipes - 1) / kSpPipes;
            const int grid = need < n_sm ? need : n_sm;
            if (pl->fp32_fft)
                k_fe_fft<float><<<grid, kSpThreads, sizeof(SpSmem<float>), st>>>(wav, tiles, pre_int[n], tb, fp, stat, pdb);
            else
                k_fe_fft<double><<<grid, kSpThreads, sizeof(SpSmem<double>), st>>>(wav, tiles, pre_int[n], tb, fp, stat, pdb);
            SC_LAUNCHED();
            rg.tile_prefix = at<int32_t>(pl, o_pmel);
            k_fe_mel<<<pre_mel[n], kMelThreads, fe_mel_smem_bytes(pl->prm.n_mels), st>>>(rg, stat, pdb, mel_raw, pl->ws_brec,
                                                                                         pl->ws_wt, pl->ws_mel);
            SC_LAUNCHED();
        } else if (ws && pre_int[n] > 0) {
            // warp-specialised persistent kernel (fe_ws.cuh): one CTA per SM
            static bool ws_attr = false;
            const size_t mel_s_bytes = sizeof(float) * 2 * kWsFrames * (pl->prm.n_mels | 1);
            if (!ws_attr) {
                SC_CUDA(cudaFuncSetAttribute(k_fe_pass_a_ws<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(sizeof(WsSmem<float>) + sizeof(float) * 2 * kWsFrames * (kMaxMels | 1))));
                SC_CUDA(cudaFuncSetAttribute(k_fe_pass_a_ws<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(sizeof(WsSmem<double>) + sizeof(float) * 2 * kWsFrames * (kMaxMels | 1))));
                ws_attr = true;
            }
            WsTile* tiles = reinterpret_cast<WsTile*>(wb + w_tiles);
            const int grid = pre_int[n] < n_sm ? pre_int[n] : n_sm;
            if (pl->fp32_fft)
                k_fe_pass_a_ws<float><<<grid, kWsThreads, sizeof(WsSmem<float>) + mel_s_bytes, st>>>(
                    wav, tiles, pre_int[n], tb, fp, stat, pdb, mel_raw, pl->ws_brec, pl->ws_wt, pl->ws_mel);
            else
                k_fe_pass_a_ws<double><<<grid, kWsThreads, sizeof(WsSmem<double>) + mel_s_bytes, st>>>(
                    wav, tiles, pre_int[n], tb, fp, stat, pdb, mel_raw, pl->ws_brec, pl->ws_wt, pl->ws_mel);
            SC_LAUNCHED();
        } else if (pre_int[n] > 0) {
            // persistent, warp-specialised, cp.async ring
            rg.tile_prefix = at<int32_t>(pl, o_pint);
            const int per_sm = pl->fp32_fft ? 3 : 2;
            const int grid = pre_int[n] < n_sm * per_sm ? pre_int[n] : n_sm * per_sm;
            if (pl->fp32_fft)
                k_fe_pass_a_persist<float, kPU><<<grid, kPU * kUnitThreads + 32, sizeof(FeSmemP<float, kPU>) + mel_bytes, st>>>(
                    wav, rg, pre_int[n], tb, fp, stat, pdb, mel_raw);
            else
                k_fe_pass_a_persist<double, kPU><<<grid, kPU * kUnitThreads + 32, sizeof(FeSmemP<double, kPU>) + mel_bytes, st>>>(
                    wav, rg, pre_int[n], tb, fp, stat, pdb, mel_raw);
            SC_LAUNCHED();
        }
        // utterances too short for the persistent kernel: one tile per CTA
        if (pre_a[n] > 0) {
            rg.tile_prefix = at<int32_t>(pl, o_pa);
            if (pl->fp32_fft)
                k_fe_pass_a<float, kPU><<<pre_a[n], kPU * kUnitThreads, sizeof(FeSmemA<float, kPU>) + mel_bytes, st>>>(wav, rg, tb, fp, stat, pdb, mel_raw);
            else
                k_fe_pass_a<double, kPU><<<pre_a[n], kPU * kUnitThreads, sizeof(FeSmemA<double, kPU>) + mel_bytes, st>>>(wav, rg, tb, fp, stat, pdb, mel_raw);
            SC_LAUNCHED();
        }
        rg.int_first = nullptr; rg.int_count = nullptr;
    } else {
        GenTables gt{pl->g_fe_win, pl->g_wn_d, pl->prm.n_fft, pl->n_bins, pl->prm.hop_length};
        const size_t smem = gen_fe_smem_bytes(pl->prm.n_fft, pl->prm.hop_length, pl->prm.n_mels);
        SC_CUDA(cudaFuncSetAttribute(k_gen_fe_pass_a, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_gen_fe_pass_a<<<pre_a[n], kGenThreads, smem, st>>>(wav, rg, gt, tb, fp, stat, pdb, mel_raw);
        SC_LAUNCHED();
    }
    if (pl->profile) if (int rc = prof_event(pl, 4 * group + 2, st)) return rc;
    rg.tile_prefix = at<int32_t>(pl, o_pb);
    // vector form of pass B: every row start must be 16-byte aligned in all four buffers
    bool vec_b = pl->use_b2 && pl->prm.n_mels % 8 == 0 && pl->prm.n_mfcc % 4 == 0 &&
                 ((reinterpret_cast<uintptr_t>(pdb) | reinterpret_cast<uintptr_t>(mel) | reinterpret_cast<uintptr_t>(mfcc)) & 15) == 0;
    for (int u = 0; u < n && vec_b; ++u) vec_b = (fo[u] & 3) == 0;
    if (vec_b && b3) {
        B3Tile* btiles = reinterpret_cast<B3Tile*>(wb + w_b3);
        k_fe_c00<<<(n + 3) / 4, 128, 0, st>>>(rg, tb, fp, stat, mel_raw);
        SC_LAUNCHED();
        {
            static int n_sm_b = 0;
            if (!n_sm_b) SC_CUDA(cudaDeviceGetAttribute(&n_sm_b, cudaDevAttrMultiProcessorCount, pl->device));
            const int grid = pre_b3[n] < 3 * n_sm_b ? pre_b3[n] : 3 * n_sm_b;
            k_fe_pass_b3<<<grid, kB3Threads, 0, st>>>(btiles, pre_b3[n], tb, fp, stat, mel_raw, pdb, mel, mfcc);
        }
        SC_LAUNCHED();
    } else if (vec_b) {
        k_fe_c00<<<(n + 3) / 4, 128, 0, st>>>(rg, tb, fp, stat, mel_raw);
        SC_LAUNCHED();
        const size_t smem = fb2_layout(pl->prm.n_mels, pl->prm.n_mfcc).bytes;
        SC_CUDA(cudaFuncSetAttribute(k_fe_pass_b2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_fe_pass_b2<<<pre_b[n], kFbThreads, smem, st>>>(rg, tb, fp, stat, mel_raw, pdb, mel, mfcc, pl->n_bins);
        SC_LAUNCHED();
    } else {
        const size_t smem = fb_layout(pl->prm.n_mels, pl->prm.n_mfcc).bytes;
        SC_CUDA(cudaFuncSetAttribute(k_fe_pass_b, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_fe_pass_b<<<pre_b[n], kFbThreads, smem, st>>>(rg, tb, fp, stat, mel_raw, pdb, mel, mfcc, pl->n_bins);
        SC_LAUNCHED();
    }
    if (pl->profile) if (int rc = prof_event(pl, 4 * group + 3, st)) return rc;
    return SC_OK;
}

extern "C" int sc_frontend_batch(sc_plan* pl, const float* wav, const int64_t* soff, const int64_t* slen_in,
                                 int32_t n, float* mfcc, float* mel, float* pdb, const int64_t* foff, void* stream) {
    if (!pl || !wav || !soff || !mfcc || !mel || !pdb || !foff) return fail(SC_ERR_INVALID, "sc_frontend_batch: null argument");
    if (n <= 0) return SC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int hop = pl->prm.hop_length;
    std::vector<int64_t> slen(n);
    int64_t max_row = 0, heap_total = 0;
    for (int u = 0; u < n; ++u) {
        slen[u] = slen_in ? slen_in[u] : soff[u + 1] - soff[u];
        if (slen[u] < 1) return fail(SC_ERR_INVALID, "sc_frontend_batch: empty utterance");
        const int64_t T = 1 + slen[u] / hop;
        if (T > INT32_MAX / 4) return fail(SC_ERR_INVALID, "sc_frontend_batch: utterance too long");
        if (pl->prm.calc_mfcc_derivative && T < 2)
            return fail(SC_ERR_INVALID, "calc_mfcc_derivate needs at least 2 frames (len >= hop_length)");
        if (foff[u] + T > max_row) max_row = foff[u] + T;
        heap_total += int64_t(2) << abs_depth(slen[u]);
    }
    // one allocation that fits every group (frontend_range never has to grow it mid-batch)
    const size_t bound = ((sizeof(UttStat) * n + 255) & ~size_t(255)) + ((sizeof(float) * (size_t)heap_total + 255) & ~size_t(255)) +
                         256 + sizeof(float) * (size_t)max_row * pl->prm.n_mels +
                         1024 + ((size_t)max_row / 20 + 2 * (size_t)n + 16) * (sizeof(WsTile) + sizeof(B3Tile)) +
                         256 + sizeof(AbsRec) * (size_t)(heap_total / 2);   // tile and sub-tree records
    if (int rc = pl->work.ensure(bound)) return rc;
    // Groups of consecutive utterances whose raw dB intermediates (1 124 B/frame) stay L2-resident between
    // pass A and pass B: pass B then re-reads from L2 and the raw values are overwritten before they reach DRAM.
    int group = 0;
    for (int u0 = 0; u0 < n;) {
        int u1 = u0;
        int64_t frames = 0;
        while (u1 < n && (u1 == u0 || frames + 1 + slen[u1] / hop <= pl->fe_group_frames)) {
            frames += 1 + slen[u1] / hop;
            ++u1;
        }
        if (int rc = frontend_range(pl, wav, soff + u0, slen.data() + u0, u1 - u0, mfcc, mel, pdb, foff + u0, st, group)) return rc;
        u0 = u1;
        ++group;
    }
    return SC_OK;
}

// np.abs(y).mean() per utterance, bit-identical to NumPy's float32 pairwise summation (:126)
extern "C" int sc_mean_abs_batch(sc_plan* pl, const float* wav, const int64_t* soff, const int64_t* slen_in, int32_t n,
                                 float* mean_out, void* stream) {
    if (!pl || !wav || !soff || !mean_out) return fail(SC_ERR_INVALID, "sc_mean_abs_batch: null argument");
    if (n <= 0) return SC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<int64_t> so(soff, soff + n), slen(n), heap_off(n + 1), fo(n, 0);
    std::vector<int32_t> pre(n + 1), fc(n, 0);
    pre[0] = 0; heap_off[0] = 0;
    for (int u = 0; u < n; ++u) {
        slen[u] = slen_in ? slen_in[u] : soff[u + 1] - soff[u];
        if (slen[u] < 1) return fail(SC_ERR_INVALID, "sc_mean_abs_batch: empty utterance");
        const int64_t t = pre[u] + (int64_t(1) << abs_depth(slen[u]));
        if (t > INT32_MAX) return fail(SC_ERR_INVALID, "sc_mean_abs_batch: batch too large");
        pre[u + 1] = (int32_t)t;
        heap_off[u + 1] = heap_off[u] + (int64_t(2) << abs_depth(slen[u]));
    }
    Blob b;
    const size_t o_so = b.add(so), o_sl = b.add(slen), o_fo = b.add(fo), o_fc = b.add(fc), o_p = b.add(pre), o_h = b.add(heap_off);
    const size_t o_arec = b.add(abs_recs(pre, n));
    if (int rc = upload_blob(pl, b, st)) return rc;
    const size_t w_heap = (sizeof(UttStat) * n + 255) & ~size_t(255);
    if (int rc = pl->work.ensure(w_heap + sizeof(float) * (size_t)heap_off[n])) return rc;
    unsigned char* wb = static_cast<unsigned char*>(pl->work.p);
    Ragged rg;
    rg.sample_off = at<int64_t>(pl, o_so); rg.sample_len = at<int64_t>(pl, o_sl);
    rg.frame_off = at<int64_t>(pl, o_fo); rg.frame_cnt = at<int32_t>(pl, o_fc);
    rg.tile_prefix = at<int32_t>(pl, o_p); rg.n_utts = n;
    float* heap = reinterpret_cast<float*>(wb + w_heap);
    if (abs_variant() >= 3) k_abs_pairwise3<<<pre[n], kAbs3Threads, 0, st>>>(wav, rg, at<int2>(pl, o_arec), at<int64_t>(pl, o_h), heap);
    else if (abs_variant() == 2) k_abs_pairwise2<<<pre[n], kAbs2Threads, 0, st>>>(wav, rg, at<int2>(pl, o_arec), at<int64_t>(pl, o_h), heap);
    else k_abs_pairwise<<<pre[n], kAbsThreads, 0, st>>>(wav, rg, at<int64_t>(pl, o_h), heap);
    SC_LAUNCHED();
    k_gain_finalize<<<(n + 3) / 4, 128, 0, st>>>(rg, at<int64_t>(pl, o_h), heap, reinterpret_cast<UttStat*>(wb), 1.0, 1, mean_out);
    SC_LAUNCHED();
    return SC_OK;
}

// --------------------------------------------------------------------------- frame labels
extern "C" int sc_phn_target_batch(sc_plan* pl, const int32_t* phn_start, const int32_t* phn_end, const int64_t* phn_off,
                                   const int64_t* slen, int32_t n, int32_t hop, int32_t win, int32_t* out_index,
                                   const int64_t* foff, void* stream) {
    if (!pl || !phn_start || !phn_end || !phn_off || !slen || !out_index || !foff)
        return fail(SC_ERR_INVALID, "sc_phn_target_batch: null argument");
    if (n <= 0) return SC_OK;
    if (hop < 1 || win < 1) return fail(SC_ERR_INVALID, "sc_phn_target_batch: hop_length and win_length must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<int64_t> po(phn_off, phn_off + n + 1), fo(foff, foff + n);
    std::vector<int32_t> fc(n);
    int64_t max_t = 0;
    for (int u = 0; u < n; ++u) {
        if (po[u + 1] - po[u] < 1) return fail(SC_ERR_INVALID, "sc_phn_target_batch: every utterance needs at least one interval");
        if (slen[u] < 0) return fail(SC_ERR_INVALID, "sc_phn_target_batch: negative length");
        const int64_t T = 1 + slen[u] / hop;
        if (T > INT32_MAX) return fail(SC_ERR_INVALID, "sc_phn_target_batch: utterance too long");
        fc[u] = (int32_t)T;
        if (T > max_t) max_t = T;
    }
    Blob b;
    const size_t o_po = b.add(po), o_fo = b.add(fo), o_fc = b.add(fc);
    if (int rc = upload_blob(pl, b, st)) return rc;
    PhnBatch pb;
    pb.start = phn_start; pb.end = phn_end;
    pb.phn_off = at<int64_t>(pl, o_po); pb.frame_off = at<int64_t>(pl, o_fo); pb.frame_cnt = at<int32_t>(pl, o_fc);
    pb.n_utts = n; pb.hop = hop; pb.win = win;
    for (int u0 = 0; u0 < n; u0 += 32768) {
        const int ny = n - u0 < 32768 ? n - u0 : 32768;
        k_phn_target<<<dim3((unsigned)((max_t + 255) / 256), (unsigned)ny), 256, 0, st>>>(pb, u0, out_index);
        SC_LAUNCHED();
    }
    return SC_OK;
}

// --------------------------------------------------------------------------- pre-emphasis
extern "C" int sc_preemphasis(const float* wav, int64_t n, double coeff, double* out, void* stream) {
    if (!wav || !out || n < 0) return fail(SC_ERR_INVALID, "sc_preemphasis: bad argument");
    if (n == 0) return SC_OK;
    k_preemph<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(wav, n, coeff, out);
    SC_LAUNCHED();
    return SC_OK;
}

static int iir_run(DevBuf& work, DescStage& ds, const float* wav, const std::vector<int64_t>& off,
                   const std::vector<int64_t>& len, double coeff, bool renorm, double target, double* out,
                   cudaStream_t st) {
    const int n = (int)off.size();
    std::vector<WavJob> jobs(n);
    std::vector<int32_t> prefix(n + 1);
    prefix[0] = 0;
    int64_t chunks = 0;
    for (int u = 0; u < n; ++u) {
        jobs[u].off = off[u]; jobs[u].len = len[u];
        jobs[u].tile0 = prefix[u]; jobs[u].chunk0 = (int32_t)chunks;
        const int64_t c = (len[u] + kIirChunk - 1) / kIirChunk;
        chunks += ((c + kIirBlock - 1) / kIirBlock) * kIirBlock;
        const int64_t t = prefix[u] + (c + kIirBlock - 1) / kIirBlock;
        if (t > INT32_MAX || chunks > INT32_MAX) return fail(SC_ERR_INVALID, "signal too long");
        prefix[u + 1] = (int32_t)t;
    }
    Blob b;
    const size_t o_jobs = b.add(jobs), o_pre = b.add(prefix);
    if (int rc = upload_blob(ds, b, st)) return rc;
    const size_t w_part = (sizeof(double) * (size_t)chunks + 255) & ~size_t(255);
    if (int rc = work.ensure(w_part + sizeof(double) * prefix[n])) return rc;
    double* chunk_end = static_cast<double*>(work.p);
    double* abs_part = reinterpret_cast<double*>(static_cast<unsigned char*>(work.p) + w_part);
    const WavJob* djobs = at<WavJob>(ds, o_jobs);
    const int32_t* dpre = at<int32_t>(ds, o_pre);
    if (prefix[n] == 0) return SC_OK;
    k_iir_local<<<prefix[n], kIirBlock, 0, st>>>(wav, djobs, n, dpre, coeff, chunk_end);
    SC_LAUNCHED();
    k_iir_carry<<<(n + 31) / 32, 32, 0, st>>>(djobs, n, coeff, chunk_end);
    SC_LAUNCHED();
    k_iir_apply<<<prefix[n], kIirBlock, 0, st>>>(wav, djobs, n, dpre, coeff, chunk_end, out, abs_part);
    SC_LAUNCHED();
    if (renorm) {
        k_renorm<<<prefix[n], 256, 0, st>>>(djobs, n, dpre, abs_part, target, out);
        SC_LAUNCHED();
    }
    return SC_OK;
}

extern "C" int sc_inv_preemphasis(const float* wav, int64_t n, double coeff, double* out, void* stream) {
    if (!wav || !out || n < 0) return fail(SC_ERR_INVALID, "sc_inv_preemphasis: bad argument");
    if (n == 0) return SC_OK;
    static thread_local DevBuf work;
    static thread_local DescStage ds;
    return iir_run(work, ds, wav, {0}, {n}, coeff, false, 0.0, out, (cudaStream_t)stream);
}

extern "C" int sc_deemph_renorm_batch(sc_plan* pl, const float* wav, const int64_t* soff, const int64_t* slen_in,
                                      int32_t n, double coeff, double target, double* out, void* stream) {
    if (!pl || !wav || !soff || !out) return fail(SC_ERR_INVALID, "sc_deemph_renorm_batch: null argument");
    if (n <= 0) return SC_OK;
    std::vector<int64_t> off(soff, soff + n), len(n);
    for (int u = 0; u < n; ++u) {
        len[u] = slen_in ? slen_in[u] : soff[u + 1] - soff[u];
        if (len[u] < 1) return fail(SC_ERR_INVALID, "sc_deemph_renorm_batch: empty signal");
    }
    return iir_run(pl->work2, pl->ds, wav, off, len, coeff, true, target, out, (cudaStream_t)stream);
}

// --------------------------------------------------------------------------- power -> amp
extern "C" int sc_power_to_amp_batch(sc_plan* pl, const float* p, const int64_t* foff, const int64_t* fcnt_in,
                                     int32_t n, double norm, double realse, float* amp, void* stream) {
    if (!pl || !p || !foff || !amp) return fail(SC_ERR_INVALID, "sc_power_to_amp_batch: null argument");
    if (n <= 0) return SC_OK;
    if (norm == 0.0) return fail(SC_ERR_INVALID, "P_dB_norm_factor must be non-zero");
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<P2aJob> jobs(n);
    std::vector<int32_t> prefix(n + 1);
    prefix[0] = 0;
    for (int u = 0; u < n; ++u) {
        jobs[u].row0 = foff[u];
        jobs[u].rows = fcnt_in ? fcnt_in[u] : foff[u + 1] - foff[u];
        if (jobs[u].rows < 1) return fail(SC_ERR_INVALID, "sc_power_to_amp_batch: empty spectrogram");
        jobs[u].tile0 = prefix[u]; jobs[u].pad = 0;
        const int64_t t = prefix[u] + (jobs[u].rows * pl->n_bins + kP2aChunk - 1) / kP2aChunk;
        if (t > INT32_MAX) return fail(SC_ERR_INVALID, "batch too large");
        prefix[u + 1] = (int32_t)t;
    }
    Blob b;
    const size_t o_j = b.add(jobs), o_p = b.add(prefix);
    if (int rc = upload_blob(pl, b, st)) return rc;
    const size_t w_scale = (sizeof(double) * 2 * (size_t)prefix[n] + 255) & ~size_t(255);
    if (int rc = pl->work2.ensure(w_scale + sizeof(float) * n)) return rc;
    double* partial = static_cast<double*>(pl->work2.p);
    float* scale = reinterpret_cast<float*>(static_cast<unsigned char*>(pl->work2.p) + w_scale);
    const int use_realse = realse != 1.0;
    if (use_realse) {
        k_p2a_partial<<<prefix[n], 256, 0, st>>>(p, at<P2aJob>(pl, o_j), n, at<int32_t>(pl, o_p), pl->n_bins,
                                                 (float)realse, partial);
        SC_LAUNCHED();
        k_p2a_scale<<<(n + 3) / 4, 128, 0, st>>>(at<P2aJob>(pl, o_j), n, at<int32_t>(pl, o_p), partial, scale);
        SC_LAUNCHED();
    }
    k_p2a_apply<<<prefix[n], 256, 0, st>>>(p, at<P2aJob>(pl, o_j), n, at<int32_t>(pl, o_p), pl->n_bins,
                                           (float)realse, use_realse, scale, (float)(1.0 / norm), amp);
    SC_LAUNCHED();
    return SC_OK;
}

// --------------------------------------------------------------------------- Griffin-Lim
static GlTables gl_tables(const sc_plan* pl) {
    GlTables t;
    t.w400 = pl->w400; t.win_half = pl->gl_win_half; t.win_inv = pl->gl_win_inv;
    t.win_sq = pl->gl_win_sq; t.inv_wss = pl->gl_inv_wss; t.zero_row = pl->zero_row;
    return t;
}

static int gl_launch(sc_plan* pl, bool init, const GlJob* jobs, int n, const int32_t* prefix, int n_tiles,
                     const float* amp, const float* phase0, const float* wav_in, float* wav_out, cudaStream_t st) {
    if (n_tiles == 0) return SC_OK;
    if (pl->fast) {
        static bool attr_set = false;
        if (!attr_set) {
            SC_CUDA(cudaFuncSetAttribute(k_gl_iter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GlSmem)));
            SC_CUDA(cudaFuncSetAttribute(k_gl_iter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GlSmem)));
            attr_set = true;
        }
        const GlTables tb = gl_tables(pl);
        if (init)
            k_gl_iter<true><<<n_tiles, kFeThreads, sizeof(GlSmem), st>>>(jobs, n, prefix, tb, amp, phase0, wav_in, wav_out);
        else
            k_gl_iter<false><<<n_tiles, kFeThreads, sizeof(GlSmem), st>>>(jobs, n, prefix, tb, amp, phase0, wav_in, wav_out);
        SC_LAUNCHED();
    } else {
        GenGlTables gt{pl->g_gl_win, pl->g_wn, pl->g_win_sq, pl->gl_inv_wss, pl->prm.n_fft, pl->n_bins, pl->prm.hop_length};
        const size_t smem = gen_gl_smem_bytes(pl->prm.n_fft, pl->prm.hop_length);
        SC_CUDA(cudaFuncSetAttribute(k_gen_gl_iter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_gen_gl_iter<<<n_tiles, kGenThreads, smem, st>>>(jobs, n, prefix, gt, amp, init ? phase0 : nullptr, wav_in, wav_out);
        SC_LAUNCHED();
    }
    return SC_OK;
}

static int64_t gl_tiles_for(const sc_plan* pl, int64_t out_first, int64_t out_count) {
    const int hop = pl->prm.hop_length, half = pl->prm.n_fft / 2;
    const int64_t out_per_tile = pl->fast ? kGlOut : gen_gl_out_per_tile(pl->prm.n_fft, hop);
    const int64_t p_first = ((out_first + half) / out_per_tile) * out_per_tile;
    const int64_t p_end = out_first + out_count + half;
    return out_count > 0 ? (p_end - p_first + out_per_tile - 1) / out_per_tile : 0;
}

extern "C" int sc_griffinlim_batch(sc_plan* pl, const float* amp, const float* phase0, const int64_t* foff,
                                   const int64_t* fcnt_in, int32_t n, int32_t n_iters, float* wav, const int64_t* soff,
                                   float* rms, void* stream) {
    if (!pl || !amp || !phase0 || !foff || !wav || !soff) return fail(SC_ERR_INVALID, "sc_griffinlim_batch: null argument");
    if (n <= 0 || n_iters <= 0) return SC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int hop = pl->prm.hop_length;
    std::vector<GlJob> jobs(n);
    std::vector<int32_t> prefix(n + 1), rprefix(n + 1);
    prefix[0] = rprefix[0] = 0;
    int64_t total = 0;
    for (int u = 0; u < n; ++u) {
        const int64_t T = fcnt_in ? fcnt_in[u] : foff[u + 1] - foff[u];
        if (T < 1 || (int64_t)hop * T + pl->prm.n_fft >= INT32_MAX)
            return fail(SC_ERR_INVALID, "sc_griffinlim_batch: bad frame count (signal must stay below 2^31 samples)");
        GlJob& j = jobs[u];
        j.amp_row0 = foff[u];
        j.wav_in_off = soff[u]; j.wav_in_first = 0; j.wav_in_count = (int64_t)hop * (T - 1);
        j.wav_out_off = soff[u]; j.out_first = 0; j.out_count = (int64_t)hop * (T - 1);
        j.f_lo = 0; j.f_cnt = (int32_t)T; j.T = (int32_t)T; j.tile0 = prefix[u];
        const int64_t t = prefix[u] + gl_tiles_for(pl, 0, j.out_count);
        const int64_t r = rprefix[u] + (j.out_count + 8191) / 8192;
        if (t > INT32_MAX) return fail(SC_ERR_INVALID, "batch too large");
        prefix[u + 1] = (int32_t)t; rprefix[u + 1] = (int32_t)r;
        if (soff[u] + j.out_count > total) total = soff[u] + j.out_count;
    }
    // (job, tile) table of the persistent iteration kernel
    std::vector<int2> tile_tab;
    if (pl->fast) {
        tile_tab.reserve(prefix[n]);
        for (int u = 0; u < n; ++u)
            for (int k = 0; k < prefix[u + 1] - prefix[u]; ++k) tile_tab.push_back(make_int2(u, k));
    }
    Blob b;
    const size_t o_j = b.add(jobs), o_p = b.add(prefix), o_r = b.add(rprefix), o_tab = b.add(tile_tab);
    if (int rc = upload_blob(pl, b, st)) return rc;
    // ping-pong partner of `wav` + rms partials
    const size_t w_part = (sizeof(float) * (size_t)total + 255) & ~size_t(255);
    if (int rc = pl->work.ensure(w_part + sizeof(double) * (size_t)rprefix[n] + 8)) return rc;
    float* other = static_cast<float*>(pl->work.p);
    double* rpart = reinterpret_cast<double*>(static_cast<unsigned char*>(pl->work.p) + w_part);
    const GlJob* dj = at<GlJob>(pl, o_j);
    const int32_t* dp = at<int32_t>(pl, o_p);
    // iteration i writes buffer (n_iters - 1 - i) & 1 ? other : wav, so the last one lands in `wav`
    auto buf = [&](int i) { return ((n_iters - 1 - i) & 1) ? other : wav; };
    if (pl->profile) { if (int rc = prof_event(pl, 0, st)) return rc; pl->pev_kind = 2; pl->pev_iters = n_iters; }
    if (int rc = gl_launch(pl, true, dj, n, dp, prefix[n], amp, phase0, nullptr, buf(0), st)) return rc;
    if (pl->profile) if (int rc = prof_event(pl, 1, st)) return rc;
    static int n_sm = 0;
    if (pl->fast && n_sm == 0) {
        SC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, pl->device));
        SC_CUDA(cudaFuncSetAttribute(k_gl_iter_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GlSmemP)));
    }
    for (int i = 1; i < n_iters; ++i) {
        if (pl->fast && prefix[n] > 0) {
            const int grid = prefix[n] < 2 * n_sm ? prefix[n] : 2 * n_sm;
            k_gl_iter_persist<<<grid, kFeThreads, sizeof(GlSmemP), st>>>(dj, at<int2>(pl, o_tab), prefix[n], gl_tables(pl), amp,
                                                                         buf(i - 1), buf(i));
            SC_LAUNCHED();
        } else if (int rc = gl_launch(pl, false, dj, n, dp, prefix[n], amp, phase0, buf(i - 1), buf(i), st)) return rc;
        if (rms && rprefix[n] > 0) {
            k_rms_delta_partial<<<rprefix[n], 256, 0, st>>>(buf(i - 1), buf(i), dj, n, at<int32_t>(pl, o_r), rpart);
            SC_LAUNCHED();
            k_rms_delta_final<<<(n + 3) / 4, 128, 0, st>>>(dj, n, at<int32_t>(pl, o_r), rpart, rms, n_iters, i);
            SC_LAUNCHED();
        }
    }
    if (pl->profile) if (int rc = prof_event(pl, 2, st)) return rc;
    return SC_OK;
}

extern "C" int sc_griffinlim_chunk_step(sc_plan* pl, const float* amp, const float* phase0, int64_t first_frame,
                                        int64_t n_local, int64_t n_total, const float* wav_in, int64_t wav_first,
                                        int64_t wav_count, float* wav_out, int64_t out_first, int64_t out_count,
                                        void* stream) {
    if (!pl || !amp || !wav_out) return fail(SC_ERR_INVALID, "sc_griffinlim_chunk_step: null argument");
    if (!phase0 && !wav_in) return fail(SC_ERR_INVALID, "sc_griffinlim_chunk_step: need phase0 or wav_in");
    if (n_total < 1 || (int64_t)pl->prm.hop_length * n_total + pl->prm.n_fft >= INT32_MAX || n_local < 0 || first_frame < 0 || first_frame + n_local > n_total)
        return fail(SC_ERR_INVALID, "sc_griffinlim_chunk_step: bad frame range");
    const int64_t Lw = (int64_t)pl->prm.hop_length * (n_total - 1);
    if (out_first < 0 || out_count < 0 || out_first + out_count > Lw)
        return fail(SC_ERR_INVALID, "sc_griffinlim_chunk_step: bad output range");
    if (out_count == 0) return SC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    GlJob j;
    j.amp_row0 = 0; j.wav_in_off = 0; j.wav_in_first = wav_first; j.wav_in_count = wav_count;
    j.wav_out_off = 0; j.out_first = out_first; j.out_count = out_count;
    j.f_lo = (int32_t)first_frame; j.f_cnt = (int32_t)n_local; j.T = (int32_t)n_total; j.tile0 = 0;
    const int64_t tiles = gl_tiles_for(pl, out_first, out_count);
    if (tiles > INT32_MAX) return fail(SC_ERR_INVALID, "chunk too large");
    int32_t prefix[2] = {0, (int32_t)tiles};
    Blob b;
    const size_t o_j = b.add(&j, 1), o_p = b.add(prefix, 2);
    if (int rc = upload_blob(pl, b, st)) return rc;
    if (pl->fast && !phase0 && tiles > 0) {
        // same persistent iteration kernel as sc_griffinlim_batch (one job: no tile table needed)
        static int n_sm = 0;
        if (n_sm == 0) {
            SC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, pl->device));
            SC_CUDA(cudaFuncSetAttribute(k_gl_iter_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GlSmemP)));
        }
        const int grid = tiles < 2 * n_sm ? (int)tiles : 2 * n_sm;
        k_gl_iter_persist<<<grid, kFeThreads, sizeof(GlSmemP), st>>>(at<GlJob>(pl, o_j), nullptr, (int)tiles, gl_tables(pl), amp,
                                                                     wav_in, wav_out);
        SC_LAUNCHED();
        return SC_OK;
    }
    return gl_launch(pl, phase0 != nullptr, at<GlJob>(pl, o_j), 1, at<int32_t>(pl, o_p), (int)tiles, amp, phase0,
                     wav_in, wav_out, st);
}

// --------------------------------------------------------------------------------- helpers
extern "C" int sc_transpose_to_f32(const void* src, int32_t is_f64, int64_t rows, int64_t cols, float* dst, void* stream) {
    if (!src || !dst || rows < 0 || cols < 0) return fail(SC_ERR_INVALID, "sc_transpose_to_f32: bad argument");
    if (rows == 0 || cols == 0) return SC_OK;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    if (grid.y > 65535) return fail(SC_ERR_INVALID, "sc_transpose_to_f32: too many rows");
    if (is_f64)
        k_transpose<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)src, rows, cols, dst);
    else
        k_transpose<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src, rows, cols, dst);
    SC_LAUNCHED();
    return SC_OK;
}

extern "C" int sc_profile_enable(sc_plan* pl, int32_t on) {
    if (!pl) return fail(SC_ERR_INVALID, "sc_profile_enable: null plan");
    pl->profile = on != 0;
    pl->pev_kind = 0;
    return SC_OK;
}

extern "C" int sc_profile_read(sc_plan* pl, double* ms_out) {
    if (!pl || !ms_out) return fail(SC_ERR_INVALID, "sc_profile_read: null argument");
    for (int i = 0; i < 4; ++i) ms_out[i] = 0.0;
    if (!pl->profile || pl->pev_kind == 0) return fail(SC_ERR_INVALID, "sc_profile_read: nothing recorded");
    if (pl->pev_kind == 1) {
        SC_CUDA(cudaEventSynchronize(pl->pev[4 * (pl->pev_groups - 1) + 3]));
        for (int g = 0; g < pl->pev_groups; ++g)
            for (int i = 0; i < 3; ++i) {
                float ms = 0.f;
                SC_CUDA(cudaEventElapsedTime(&ms, pl->pev[4 * g + i], pl->pev[4 * g + i + 1]));
                ms_out[i] += ms;
            }
        ms_out[3] = (double)pl->pev_groups;
    } else {
        SC_CUDA(cudaEventSynchronize(pl->pev[2]));
        for (int i = 0; i < 2; ++i) {
            float ms = 0.f;
            SC_CUDA(cudaEventElapsedTime(&ms, pl->pev[i], pl->pev[i + 1]));
            ms_out[i] = ms;
        }
        ms_out[3] = (double)pl->pev_iters;
    }
    return SC_OK;
}

extern "C" int64_t sc_launch_count(void) { return g_launches.load(); }
extern "C" void sc_launch_count_reset(void) { g_launches.store(0); }
extern "C" const char* sc_last_error(void) { return g_err.c_str(); }
extern "C" const char* sc_version(void) { return "speechdsp-b200 0.1 (sm_100a)"; }
