// C ABI (include/speechdsp.h) + plan / host-side tables for the speech-cloner DSP hot path.
// Single translation unit: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -shared.
#include "../../include/speechdsp.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fe_kernels.cuh"
#include "fe_ws.cuh"
#include "gl_kernels.cuh"
#include "generic_kernels.cuh"
#include "phn_kernels.cuh"
#include "sampler_kernels.cuh"

using namespace scdsp;

// ------------------------------------------------------------------------------------- errors
static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};
static std::atomic<int64_t> g_allocs{0};      // cudaMalloc / cudaMallocHost calls made by this library (sc_alloc_count)

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
// Grow-only device buffer.  sc_plan_reserve() sizes every buffer of a plan up front, so that compute calls within
// the reserved bounds never allocate; a call that exceeds them still works (it grows the buffer, which synchronises).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        g_allocs.fetch_add(1, std::memory_order_relaxed);
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

// Descriptor staging: a ring of pinned host slots -> one async copy into the (single) device buffer.  The device
// buffer is reused in stream order (a plan is single-stream, see speechdsp.h); a pinned slot is reused only after the
// copy that read it has completed, so the host runs up to kSlots calls ahead of the GPU without blocking.
struct DescStage {
    static constexpr int kSlots = 4;
    DevBuf desc;
    unsigned char* pinned[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    size_t pinned_cap[kSlots] = {0, 0, 0, 0};
    cudaEvent_t done[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    bool pending[kSlots] = {false, false, false, false};
    int next = 0;
    int reserve(size_t bytes) {
        for (int s = 0; s < kSlots; ++s) {
            if (!done[s]) SC_CUDA(cudaEventCreateWithFlags(&done[s], cudaEventDisableTiming));
            if (bytes > pinned_cap[s]) {
                if (pending[s]) { SC_CUDA(cudaEventSynchronize(done[s])); pending[s] = false; }
                if (pinned[s]) cudaFreeHost(pinned[s]);
                pinned[s] = nullptr; pinned_cap[s] = 0;
                SC_CUDA(cudaMallocHost((void**)&pinned[s], bytes));
                g_allocs.fetch_add(1, std::memory_order_relaxed);
                pinned_cap[s] = bytes;
            }
        }
        return desc.ensure(bytes);
    }
    void release() {
        for (int s = 0; s < kSlots; ++s) {
            if (pending[s] && done[s]) cudaEventSynchronize(done[s]);
            if (done[s]) cudaEventDestroy(done[s]);
            if (pinned[s]) cudaFreeHost(pinned[s]);
            pinned[s] = nullptr; pinned_cap[s] = 0; done[s] = nullptr; pending[s] = false;
        }
        desc.release();
    }
};

// layout-dependent tables of the front-end (descriptor arrays + the tile / sub-tree records k_fe_setup writes), kept
// across calls: a call with the same ragged layout as the previous one re-uses them (no host vectors, no upload, no
// setup launch) - the fixed-shape batches of a dataset sweep or a serving loop hit this every time
struct FeLayoutCache {
    bool valid = false;
    std::vector<int64_t> key;             // n, then sample offsets, lengths, frame offsets
    size_t o_so = 0, o_sl = 0, o_fo = 0, o_fc = 0, o_pabs = 0, o_pa = 0, o_pb = 0, o_heap = 0, o_pint = 0, o_if = 0, o_ic = 0, o_pb3 = 0;
    int32_t n_abs = 0, n_a = 0, n_b = 0, n_b3 = 0, n_int = 0;
    int64_t heap_total = 0, max_row = 0;
    size_t t_ws = 0, t_b3 = 0, t_abs = 0;  // offsets inside fe_tab
    bool vec_rows = false;                // every utterance starts on a 4-row boundary
};

struct sc_plan {
    sc_params prm{};
    int n_bins = 0;
    bool fast = false;
    int device = 0;
    int n_sm = 0;
    // single-thread / single-stream contract (speechdsp.h): concurrent use is refused, a call on another stream is
    // ordered after the previous call through `last_done`
    std::atomic_flag busy = ATOMIC_FLAG_INIT;
    cudaEvent_t last_done = nullptr;
    cudaStream_t last_stream = nullptr;
    bool has_last = false;
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
    // workspaces
    DescStage ds;                        // descriptors of every call but the front-end
    DescStage ds_fe;                     // front-end descriptors (cached across calls, see FeLayoutCache)
    DevBuf work, work2, fe_tab;
    DevBuf status;                       // one int32: bit 0 = an utterance with mean|y| == 0 or a non-finite gain
    FeLayoutCache fe_cache;
    // optional per-kernel timing (bench.py's roofline): events recorded on the caller's stream
    bool profile = false;
    std::vector<cudaEvent_t> pev;   // 4 events for a front-end call, or 3 for a Griffin-Lim call
    int pev_kind = 0;        // 1 = front-end (gain | pass A | pass B), 2 = Griffin-Lim (init | iterations)
    int pev_iters = 0;
    WsMelParam ws_mel{};     // band / bin ranges of the epilogue warps of k_fe_pass_a_ws
    const int4* ws_brec = nullptr;  // per-band (first tap, float4 blocks, weight offset)
    const float* ws_wt = nullptr;   // padded filterbank weights
    bool use_b3 = false;     // compile-time specialised pass B (80 mels, 40 MFCC, 201 bins)
    bool use_ws = true;      // warp-specialised pass A (always, unless the filterbank does not fit its tables)
};

// Entry-point guard: one thread at a time per plan, plan used on the device it was created on, calls on different
// streams ordered one after the other.
struct PlanGuard {
    sc_plan* pl;
    bool ok;
    explicit PlanGuard(sc_plan* p) : pl(p), ok(!p->busy.test_and_set(std::memory_order_acquire)) {}
    ~PlanGuard() { if (ok) pl->busy.clear(std::memory_order_release); }
};
static int plan_enter(sc_plan* pl, const PlanGuard& g, cudaStream_t st, const char* who) {
    if (!g.ok) return fail(SC_ERR_INVALID, std::string(who) + ": the plan is in use by another thread (a plan is single-threaded: create one per thread / stream)");
    int dev = -1;
    SC_CUDA(cudaGetDevice(&dev));
    if (dev != pl->device)
        return fail(SC_ERR_INVALID, std::string(who) + ": plan was created on device " + std::to_string(pl->device) +
                                        " but the current device is " + std::to_string(dev));
    if (pl->has_last && pl->last_stream != st) {
        // order this call after everything the plan has queued on the stream it was last used on
        if (cudaEventRecord(pl->last_done, pl->last_stream) == cudaSuccess) SC_CUDA(cudaStreamWaitEvent(st, pl->last_done, 0));
        else (void)cudaGetLastError();      // the old stream is gone: its work has completed
    }
    pl->last_stream = st; pl->has_last = true;
    return 0;
}
#define SC_ENTER(pl, st, who)                                        \
    PlanGuard guard_(pl);                                            \
    if (int rc_ = plan_enter(pl, guard_, st, who)) return rc_

static int upload_blob(DescStage& ds, const Blob& b, cudaStream_t st) {
    const int s = ds.next;
    ds.next = (ds.next + 1) % DescStage::kSlots;
    if (!ds.done[s]) SC_CUDA(cudaEventCreateWithFlags(&ds.done[s], cudaEventDisableTiming));
    if (ds.pending[s]) {
        SC_CUDA(cudaEventSynchronize(ds.done[s]));
        ds.pending[s] = false;
    }
    if (b.bytes.size() > ds.pinned_cap[s]) {
        if (ds.pinned[s]) cudaFreeHost(ds.pinned[s]);
        ds.pinned[s] = nullptr; ds.pinned_cap[s] = 0;
        size_t want = b.bytes.size() * 2 + 4096;
        SC_CUDA(cudaMallocHost((void**)&ds.pinned[s], want));
        g_allocs.fetch_add(1, std::memory_order_relaxed);
        ds.pinned_cap[s] = want;
    }
    if (int rc = ds.desc.ensure(b.bytes.size())) return rc;
    memcpy(ds.pinned[s], b.bytes.data(), b.bytes.size());
    SC_CUDA(cudaMemcpyAsync(ds.desc.p, ds.pinned[s], b.bytes.size(), cudaMemcpyHostToDevice, st));
    SC_CUDA(cudaEventRecord(ds.done[s], st));
    ds.pending[s] = true;
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
    const double band_cost = 20.0;                            // instructions per closed band relative to 3 per bin
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
// Per-device state of the kernels (the opt-in shared-memory size is a per-device function attribute) and the SM count
// that sizes the persistent grids.  Runs at every plan creation, so a process that drives several GPUs gets the
// attributes on each of them.
static int device_setup(sc_plan* pl) {
    constexpr int kPU = 8;
    SC_CUDA(cudaDeviceGetAttribute(&pl->n_sm, cudaDevAttrMultiProcessorCount, pl->device));
    SC_CUDA(cudaFuncSetAttribute(k_fe_pass_a<float, kPU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    SC_CUDA(cudaFuncSetAttribute(k_fe_pass_a<double, kPU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    SC_CUDA(cudaFuncSetAttribute(k_fe_pass_a_ws<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(WsSmem<float>) + sizeof(float) * 2 * kWsFrames * (kMaxMels | 1))));
    SC_CUDA(cudaFuncSetAttribute(k_fe_pass_a_ws<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(WsSmem<double>) + sizeof(float) * 2 * kWsFrames * (kMaxMels | 1))));
    SC_CUDA(cudaFuncSetAttribute(k_gl_iter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GlSmem)));
    SC_CUDA(cudaFuncSetAttribute(k_gl_iter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GlSmem)));
    SC_CUDA(cudaFuncSetAttribute(k_gl_iter_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GlSmemP)));
    // (the generic-size kernels and the non-specialised pass B set their size per launch: it depends on the plan)
    SC_CUDA(cudaEventCreateWithFlags(&pl->last_done, cudaEventDisableTiming));
    return 0;
}

extern "C" void sc_plan_destroy(sc_plan* pl);
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
    cudaGetDevice(&pl->device);
    if (int rc = device_setup(pl)) { delete pl; return rc; }

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
        const double ws_cost_block = 1.0, ws_cost_band = 3.4;
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
    pl->use_b3 = pl->n_bins == kBins && p->n_mels == kB3Mels && p->n_mfcc == kB3Mfcc;
    const size_t o_wn = b.add(wn), o_feh = b.add(fe_half), o_glh = b.add(gl_half), o_gli = b.add(gl_inv);
    const size_t o_sq = b.add(gl_sq), o_wss = b.add(inv_wss), o_mw = b.add(ms.w), o_mi = b.add(ms.istart);
    const size_t o_mc = b.add(ms.chunk), o_dcte = b.add(dct_e), o_dcto = b.add(dct_o), o_few = b.add(pl->fe_window), o_glw = b.add(gl_w);
    const size_t o_wnd = b.add(wn_d), o_fehd = b.add(fe_half_d), o_wsrec = b.add(ws_brec), o_wswt = b.add(ws_wt);
    std::vector<float> zeros(pl->n_bins, 0.f);
    const size_t o_zero = b.add(zeros);
    if (pl->tables.ensure(b.bytes.size()) || pl->status.ensure(sizeof(int32_t))) { sc_plan_destroy(pl); return SC_ERR_CUDA; }
    e = cudaMemcpy(pl->tables.p, b.bytes.data(), b.bytes.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(pl->status.p, 0, sizeof(int32_t));
    if (e != cudaSuccess) {
        sc_plan_destroy(pl);
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
    pl->ds.release(); pl->ds_fe.release();
    for (auto& e : pl->pev) if (e) cudaEventDestroy(e);
    if (pl->last_done) cudaEventDestroy(pl->last_done);
    pl->tables.release(); pl->work.release(); pl->work2.release(); pl->fe_tab.release(); pl->status.release();
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
// Worst-case sizes (bytes) of the front-end buffers for a batch of n utterances, `samples` samples and `rows` feature
// rows in total: used by sc_plan_reserve; the calls themselves ask for their exact sizes, which are never larger.
static size_t fe_heap_floats_bound(int64_t samples, int64_t n) { return (size_t)(samples / (kAbsSubtree / 4) + 4 * n + 16); }
static size_t fe_work_bytes(const sc_plan* pl, int64_t n, size_t heap_floats, int64_t rows) {
    return ((sizeof(UttStat) * (size_t)n + 255) & ~size_t(255)) + ((sizeof(float) * heap_floats + 255) & ~size_t(255)) +
           sizeof(float) * (size_t)rows * pl->prm.n_mels + 512;
}
static size_t fe_tab_bytes(int64_t n_ws, int64_t n_b3, int64_t n_abs) {
    return ((sizeof(WsTile) * (size_t)n_ws + 255) & ~size_t(255)) + ((sizeof(B3Tile) * (size_t)n_b3 + 255) & ~size_t(255)) +
           sizeof(AbsRec) * (size_t)n_abs + 256;
}
static size_t fe_desc_bytes(int64_t n) { return (size_t)n * (4 * 8 + 8 * 4) + 16 * 16 + 64; }

// Host side of one front-end call: derive the per-utterance tables for this ragged layout (or find them cached),
// make them resident, and run  setup -> |y| sums -> gains -> pass A -> MFCC[0,0] -> pass B  on `st`.
static int frontend_run(sc_plan* pl, const float* wav, const int64_t* soff, const int64_t* slen_in, int32_t n,
                        float* mfcc, float* mel, float* pdb, const int64_t* foff, cudaStream_t st) {
    const int hop = pl->prm.hop_length;
    constexpr int kPU = 8;                         // units per CTA of the one-tile pass A kernel (16 frames per tile)
    const bool ws = pl->fast && pl->use_ws;        // long utterances: warp-specialised kernel, 24-frame tiles
    const int a_frames = pl->fast ? 2 * kPU : kGenFeFrames;
    FeLayoutCache& fc = pl->fe_cache;

    // ---- cache key: the layout arrays exactly as passed
    std::vector<int64_t> key((size_t)3 * n + 1);
    key[0] = n;
    for (int u = 0; u < n; ++u) {
        key[1 + u] = soff[u];
        key[1 + n + u] = slen_in ? slen_in[u] : soff[u + 1] - soff[u];
        key[1 + 2 * n + u] = foff[u];
    }
    const bool hit = fc.valid && fc.key == key;
    if (!hit) {
        fc.valid = false;
        std::vector<int64_t> slen(key.begin() + 1 + n, key.begin() + 1 + 2 * n), so(soff, soff + n), fo(foff, foff + n);
        std::vector<int32_t> fcnt(n), pre_abs(n + 1), pre_a(n + 1), pre_b(n + 1), pre_b3(n + 1), pre_int(n + 1), ifirst(n), icount(n);
        std::vector<int64_t> heap_off(n + 1);
        int64_t max_row = 0;
        bool vec_rows = true;
        pre_abs[0] = pre_a[0] = pre_b[0] = pre_b3[0] = pre_int[0] = 0;
        heap_off[0] = 0;
        for (int u = 0; u < n; ++u) {
            if (slen[u] < 1) return fail(SC_ERR_INVALID, "sc_frontend_batch: empty utterance");
            const int64_t T = 1 + slen[u] / hop;
            if (T > INT32_MAX / 4) return fail(SC_ERR_INVALID, "sc_frontend_batch: utterance too long");
            if (pl->prm.calc_mfcc_derivative && T < 2)
                return fail(SC_ERR_INVALID, "calc_mfcc_derivate needs at least 2 frames (len >= hop_length)");
            fcnt[u] = (int32_t)T;
            if (fo[u] + T > max_row) max_row = fo[u] + T;
            vec_rows = vec_rows && (fo[u] & 3) == 0;
            const int64_t ta = pre_abs[u] + (int64_t(1) << abs_depth(slen[u]));
            heap_off[u + 1] = heap_off[u] + (int64_t(2) << abs_depth(slen[u]));
            // fast path: every tile of an utterance at least two tile spans long runs in the persistent kernel (its
            // first / last tiles gather their reflect padding); shorter utterances, whose padding could wrap more
            // than once, use the one-tile-per-CTA kernel
            const int64_t n_tiles_u = (T + a_frames - 1) / a_frames;
            int64_t n_int = 0, n_edge = n_tiles_u;
            if (ws && slen[u] >= 2 * (int64_t)kWsSpan + 16) { n_int = (T + kWsFrames - 1) / kWsFrames; n_edge = 0; }
            ifirst[u] = 0; icount[u] = (int32_t)(n_tiles_u - n_edge);
            const int64_t tb = pre_a[u] + n_edge;
            const int64_t tc = pre_b[u] + (T + kFbFrames - 1) / kFbFrames;
            if (ta > INT32_MAX || tb > INT32_MAX || tc > INT32_MAX) return fail(SC_ERR_INVALID, "sc_frontend_batch: batch too large");
            pre_abs[u + 1] = (int32_t)ta; pre_a[u + 1] = (int32_t)tb; pre_b[u + 1] = (int32_t)tc;
            pre_int[u + 1] = (int32_t)(pre_int[u] + n_int);
            pre_b3[u + 1] = pre_b3[u] + (int32_t)((T + kB3Frames - 1) / kB3Frames);
        }
        Blob b;
        fc.o_so = b.add(so); fc.o_sl = b.add(slen); fc.o_fo = b.add(fo); fc.o_fc = b.add(fcnt);
        fc.o_pabs = b.add(pre_abs); fc.o_pa = b.add(pre_a); fc.o_pb = b.add(pre_b); fc.o_heap = b.add(heap_off);
        fc.o_pint = b.add(pre_int); fc.o_if = b.add(ifirst); fc.o_ic = b.add(icount); fc.o_pb3 = b.add(pre_b3);
        if (int rc = upload_blob(pl->ds_fe, b, st)) return rc;
        fc.n_abs = pre_abs[n]; fc.n_a = pre_a[n]; fc.n_b = pre_b[n]; fc.n_b3 = pre_b3[n]; fc.n_int = pre_int[n];
        fc.heap_total = heap_off[n]; fc.max_row = max_row; fc.vec_rows = vec_rows;
        fc.t_ws = 0;
        fc.t_b3 = (sizeof(WsTile) * (size_t)fc.n_int + 255) & ~size_t(255);
        fc.t_abs = (fc.t_b3 + sizeof(B3Tile) * (size_t)fc.n_b3 + 255) & ~size_t(255);
        if (int rc = pl->fe_tab.ensure(fc.t_abs + sizeof(AbsRec) * (size_t)fc.n_abs)) return rc;
        fc.key.swap(key);
    }
    const DescStage& dd = pl->ds_fe;

    // per-call scratch: stats | |y| partial sums | raw mel dB
    const size_t w_part = (sizeof(UttStat) * (size_t)n + 255) & ~size_t(255);
    const size_t w_mel = (w_part + sizeof(float) * (size_t)fc.heap_total + 255) & ~size_t(255);
    if (int rc = pl->work.ensure(w_mel + sizeof(float) * (size_t)fc.max_row * pl->prm.n_mels)) return rc;
    unsigned char* wb = static_cast<unsigned char*>(pl->work.p);
    unsigned char* tabs = static_cast<unsigned char*>(pl->fe_tab.p);
    UttStat* stat = reinterpret_cast<UttStat*>(wb);
    float* heap = reinterpret_cast<float*>(wb + w_part);
    float* mel_raw = reinterpret_cast<float*>(wb + w_mel);
    WsTile* ws_tiles = reinterpret_cast<WsTile*>(tabs + fc.t_ws);
    B3Tile* b3_tiles = reinterpret_cast<B3Tile*>(tabs + fc.t_b3);
    AbsRec* abs_recs = reinterpret_cast<AbsRec*>(tabs + fc.t_abs);

    Ragged rg;
    rg.sample_off = at<int64_t>(dd, fc.o_so); rg.sample_len = at<int64_t>(dd, fc.o_sl);
    rg.frame_off = at<int64_t>(dd, fc.o_fo); rg.frame_cnt = at<int32_t>(dd, fc.o_fc);
    rg.n_utts = n;
    rg.tile_prefix = at<int32_t>(dd, fc.o_pabs);
    rg.int_first = nullptr; rg.int_count = nullptr;
    const FeTables tb = fe_tables(pl);
    const FeParams fp = fe_params(pl);
    const bool b3 = pl->fast && pl->use_b3;
    const int n_sm = pl->n_sm;

    if (pl->profile) { if (int rc = prof_event(pl, 0, st)) return rc; pl->pev_kind = 1; }
    // one launch writes every layout table (sub-tree records of the |y| sum, pass A tiles, pass B tiles); skipped when
    // the tables of the previous call are still valid
    if (!hit) {
        SetupArgs sa{};
        sa.pre_abs = at<int32_t>(dd, fc.o_pabs); sa.pre_ws = at<int32_t>(dd, fc.o_pint); sa.pre_b3 = at<int32_t>(dd, fc.o_pb3);
        sa.heap_off = at<int64_t>(dd, fc.o_heap);
        sa.n_abs = fc.n_abs;
        sa.n_ws = ws ? fc.n_int : 0;
        sa.ws_frames = kWsFrames;
        sa.n_b3 = b3 ? fc.n_b3 : 0;
        sa.abs_out = abs_recs; sa.ws_out = ws_tiles; sa.b3_out = b3_tiles;
        const int total = sa.n_abs + sa.n_ws + sa.n_b3;
        k_fe_setup<<<(total + 255) / 256, 256, 0, st>>>(rg, sa);
        SC_LAUNCHED();
        fc.valid = true;
    }
    if (fp.use_gain) {
        const int grid = fc.n_abs < 4 * n_sm ? fc.n_abs : 4 * n_sm;
        k_abs_pairwise4<<<grid, kAbs3Threads, 0, st>>>(wav, abs_recs, fc.n_abs, heap);
        SC_LAUNCHED();
    }
    k_gain_finalize<<<(n + 3) / 4, 128, 0, st>>>(rg, at<int64_t>(dd, fc.o_heap), heap, stat, fp.mean_abs_amp_norm, fp.use_gain,
                                                 nullptr, static_cast<int32_t*>(pl->status.p));
    SC_LAUNCHED();

    if (pl->profile) if (int rc = prof_event(pl, 1, st)) return rc;
    rg.tile_prefix = at<int32_t>(dd, fc.o_pa);
    if (pl->fast) {
        const size_t mel_bytes = sizeof(float) * 2 * kPU * (pl->prm.n_mels + 1);
        if (fc.n_int > 0) {
            // warp-specialised persistent kernel (fe_ws.cuh): one CTA per SM
            const size_t mel_s_bytes = sizeof(float) * 2 * kWsFrames * (pl->prm.n_mels | 1);
            const int grid = fc.n_int < n_sm ? fc.n_int : n_sm;
            if (pl->fp32_fft)
                k_fe_pass_a_ws<float><<<grid, kWsThreads, sizeof(WsSmem<float>) + mel_s_bytes, st>>>(
                    wav, ws_tiles, fc.n_int, tb, fp, stat, pdb, mel_raw, pl->ws_brec, pl->ws_wt, pl->ws_mel);
            else
                k_fe_pass_a_ws<double><<<grid, kWsThreads, sizeof(WsSmem<double>) + mel_s_bytes, st>>>(
                    wav, ws_tiles, fc.n_int, tb, fp, stat, pdb, mel_raw, pl->ws_brec, pl->ws_wt, pl->ws_mel);
            SC_LAUNCHED();
        }
        // utterances too short for the persistent kernel: one tile per CTA
        if (fc.n_a > 0) {
            rg.int_first = at<int32_t>(dd, fc.o_if); rg.int_count = at<int32_t>(dd, fc.o_ic);
            if (pl->fp32_fft)
                k_fe_pass_a<float, kPU><<<fc.n_a, kPU * kUnitThreads, sizeof(FeSmemA<float, kPU>) + mel_bytes, st>>>(wav, rg, tb, fp, stat, pdb, mel_raw);
            else
                k_fe_pass_a<double, kPU><<<fc.n_a, kPU * kUnitThreads, sizeof(FeSmemA<double, kPU>) + mel_bytes, st>>>(wav, rg, tb, fp, stat, pdb, mel_raw);
            SC_LAUNCHED();
            rg.int_first = nullptr; rg.int_count = nullptr;
        }
    } else {
        GenTables gt{pl->g_fe_win, pl->g_wn_d, pl->prm.n_fft, pl->n_bins, pl->prm.hop_length};
        const size_t smem = gen_fe_smem_bytes(pl->prm.n_fft, pl->prm.hop_length, pl->prm.n_mels);
        SC_CUDA(cudaFuncSetAttribute(k_gen_fe_pass_a, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_gen_fe_pass_a<<<fc.n_a, kGenThreads, smem, st>>>(wav, rg, gt, tb, fp, stat, pdb, mel_raw);
        SC_LAUNCHED();
    }
    if (pl->profile) if (int rc = prof_event(pl, 2, st)) return rc;
    rg.tile_prefix = at<int32_t>(dd, fc.o_pb);
    // vector form of pass B: every row start must be 16-byte aligned in all four buffers
    const bool vec_b = fc.vec_rows && pl->prm.n_mels % 8 == 0 && pl->prm.n_mfcc % 4 == 0 &&
                       ((reinterpret_cast<uintptr_t>(pdb) | reinterpret_cast<uintptr_t>(mel) | reinterpret_cast<uintptr_t>(mfcc)) & 15) == 0;
    if (vec_b) {
        if (b3) {
            // MFCC[0, 0] is computed inside the kernel (extra row of the folded tile)
            const int grid = fc.n_b3 < 3 * n_sm ? fc.n_b3 : 3 * n_sm;
            k_fe_pass_b3<<<grid, kB3Threads, 0, st>>>(b3_tiles, fc.n_b3, tb, fp, stat, mel_raw, pdb, mel, mfcc);
        } else {
            k_fe_c00<<<(n + 3) / 4, 128, 0, st>>>(rg, tb, fp, stat, mel_raw);
            SC_LAUNCHED();
            const size_t smem = fb2_layout(pl->prm.n_mels, pl->prm.n_mfcc).bytes;
            SC_CUDA(cudaFuncSetAttribute(k_fe_pass_b2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_fe_pass_b2<<<fc.n_b, kFbThreads, smem, st>>>(rg, tb, fp, stat, mel_raw, pdb, mel, mfcc, pl->n_bins);
        }
        SC_LAUNCHED();
    } else {
        const size_t smem = fb_layout(pl->prm.n_mels, pl->prm.n_mfcc).bytes;
        SC_CUDA(cudaFuncSetAttribute(k_fe_pass_b, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_fe_pass_b<<<fc.n_b, kFbThreads, smem, st>>>(rg, tb, fp, stat, mel_raw, pdb, mel, mfcc, pl->n_bins);
        SC_LAUNCHED();
    }
    if (pl->profile) if (int rc = prof_event(pl, 3, st)) return rc;
    return SC_OK;
}

extern "C" int sc_frontend_batch(sc_plan* pl, const float* wav, const int64_t* soff, const int64_t* slen_in,
                                 int32_t n, float* mfcc, float* mel, float* pdb, const int64_t* foff, void* stream) {
    if (!pl || !wav || !soff || !mfcc || !mel || !pdb || !foff) return fail(SC_ERR_INVALID, "sc_frontend_batch: null argument");
    if (n <= 0) return SC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    SC_ENTER(pl, st, "sc_frontend_batch");
    return frontend_run(pl, wav, soff, slen_in, n, mfcc, mel, pdb, foff, st);
}

// np.abs(y).mean() per utterance, bit-identical to NumPy's float32 pairwise summation (:126).  Same kernels as the
// gain of sc_frontend_batch (k_fe_setup records -> k_abs_pairwise4 -> k_gain_finalize).
extern "C" int sc_mean_abs_batch(sc_plan* pl, const float* wav, const int64_t* soff, const int64_t* slen_in, int32_t n,
                                 float* mean_out, void* stream) {
    if (!pl || !wav || !soff || !mean_out) return fail(SC_ERR_INVALID, "sc_mean_abs_batch: null argument");
    if (n <= 0) return SC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    SC_ENTER(pl, st, "sc_mean_abs_batch");
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
    if (int rc = upload_blob(pl, b, st)) return rc;
    const size_t w_heap = (sizeof(UttStat) * (size_t)n + 255) & ~size_t(255);
    const size_t w_rec = (w_heap + sizeof(float) * (size_t)heap_off[n] + 255) & ~size_t(255);
    if (int rc = pl->work.ensure(w_rec + sizeof(AbsRec) * (size_t)pre[n])) return rc;
    unsigned char* wb = static_cast<unsigned char*>(pl->work.p);
    Ragged rg;
    rg.sample_off = at<int64_t>(pl, o_so); rg.sample_len = at<int64_t>(pl, o_sl);
    rg.frame_off = at<int64_t>(pl, o_fo); rg.frame_cnt = at<int32_t>(pl, o_fc);
    rg.tile_prefix = at<int32_t>(pl, o_p); rg.n_utts = n;
    rg.int_first = nullptr; rg.int_count = nullptr;
    float* heap = reinterpret_cast<float*>(wb + w_heap);
    AbsRec* recs = reinterpret_cast<AbsRec*>(wb + w_rec);
    SetupArgs sa{};
    sa.pre_abs = at<int32_t>(pl, o_p); sa.pre_ws = nullptr; sa.pre_b3 = nullptr;
    sa.heap_off = at<int64_t>(pl, o_h);
    sa.n_abs = pre[n]; sa.n_ws = 0; sa.n_b3 = 0; sa.ws_frames = kWsFrames;
    sa.abs_out = recs; sa.ws_out = nullptr; sa.b3_out = nullptr;
    k_fe_setup<<<(pre[n] + 255) / 256, 256, 0, st>>>(rg, sa);
    SC_LAUNCHED();
    const int grid = pre[n] < 4 * pl->n_sm ? pre[n] : 4 * pl->n_sm;
    k_abs_pairwise4<<<grid, kAbs3Threads, 0, st>>>(wav, recs, pre[n], heap);
    SC_LAUNCHED();
    k_gain_finalize<<<(n + 3) / 4, 128, 0, st>>>(rg, at<int64_t>(pl, o_h), heap, reinterpret_cast<UttStat*>(wb), 1.0, 1, mean_out, nullptr);
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
    SC_ENTER(pl, st, "sc_phn_target_batch");
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

// --------------------------------------------------------------------------- window samplers
extern "C" int sc_window_gather(const void* const* src, void* const* dst, const int64_t* width, int32_t n_arrays,
                                int64_t n_rows_total, const int64_t* first_row, const int32_t* valid, int32_t n_windows,
                                int32_t n_timesteps, void* stream) {
    if (!src || !dst || !width || !first_row || !valid) return fail(SC_ERR_INVALID, "sc_window_gather: null argument");
    if (n_arrays < 1 || n_arrays > kGatherMaxArrays)
        return fail(SC_ERR_INVALID, "sc_window_gather: between 1 and 4 arrays per call");
    if (n_windows < 0 || n_timesteps < 1 || n_rows_total < 0) return fail(SC_ERR_INVALID, "sc_window_gather: bad size");
    GatherArgs g{};
    int64_t max_words = 0;
    for (int a = 0; a < n_arrays; ++a) {
        if (!src[a] || !dst[a] || width[a] < 1) return fail(SC_ERR_INVALID, "sc_window_gather: null array or width < 1");
        if ((reinterpret_cast<uintptr_t>(src[a]) | reinterpret_cast<uintptr_t>(dst[a])) & 3)
            return fail(SC_ERR_INVALID, "sc_window_gather: arrays must be 4-byte aligned");
        g.a[a].src = static_cast<const uint32_t*>(src[a]);
        g.a[a].dst = static_cast<uint32_t*>(dst[a]);
        g.a[a].width = width[a];
        max_words = std::max(max_words, width[a] * (int64_t)n_timesteps);
    }
    if (n_windows == 0) return SC_OK;
    const int64_t bx = std::min<int64_t>((max_words + kGatherWordsPerBlock - 1) / kGatherWordsPerBlock, 1024);
    for (int32_t w0 = 0; w0 < n_windows; w0 += 32768) {
        const int ny = n_windows - w0 < 32768 ? n_windows - w0 : 32768;
        k_window_gather<<<dim3((unsigned)bx, (unsigned)ny, (unsigned)n_arrays), kGatherThreads, 0, (cudaStream_t)stream>>>(
            g, first_row, valid, w0, n_timesteps, n_rows_total);
        SC_LAUNCHED();
    }
    return SC_OK;
}

// --------------------------------------------------------------------------- pre-emphasis
extern "C" int sc_preemphasis(const float* wav, int64_t n, double coeff, double* out, void* stream) {
    if (!wav || !out || n < 0) return fail(SC_ERR_INVALID, "sc_preemphasis: bad argument");
    if (n == 0) return SC_OK;
    k_preemph<float><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(wav, n, coeff, out);
    SC_LAUNCHED();
    return SC_OK;
}
extern "C" int sc_preemphasis_f64(const double* wav, int64_t n, double coeff, double* out, void* stream) {
    if (!wav || !out || n < 0) return fail(SC_ERR_INVALID, "sc_preemphasis_f64: bad argument");
    if (n == 0) return SC_OK;
    k_preemph<double><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(wav, n, coeff, out);
    SC_LAUNCHED();
    return SC_OK;
}

// terms of the carry series of the de-emphasis scan (gl_kernels.cuh): smallest w with |c|^(256 w) < 1e-40;
// 0 = the coefficient is too close to 1 for a windowed carry (sequential fallback)
static int iir_window(double c) {
    const double a = fabs(c);
    if (a == 0.0) return 1;
    if (a >= 1.0) return 0;
    const double w = ceil(log(1e-40) / (kIirChunk * log(a)));
    return w <= kIirMaxWin ? (int)(w < 1 ? 1 : w) : 0;
}
static double iir_chunk_factor(double c) {             // c^kIirChunk by the same repeated product as the kernels
    double cl = 1.0;
    for (int i = 0; i < kIirChunk; ++i) cl *= c;
    return cl;
}

// y[n] = x[n] + c*y[n-1] for ONE signal without a plan: scratch comes from the stream-ordered allocator and is
// returned to it before the call ends (nothing is kept between calls)
template <typename TIN>
static int inv_preemph_one(const TIN* wav, int64_t n, double coeff, double* out, cudaStream_t st) {
    const int64_t chunks = (n + kIirChunk - 1) / kIirChunk;
    const int64_t tiles = (chunks + kIirBlock - 1) / kIirBlock;
    if (tiles > INT32_MAX) return fail(SC_ERR_INVALID, "signal too long");
    const int win = iir_window(coeff);
    struct Head { WavJob job; int32_t prefix[2]; };
    Head h{};
    h.job.off = 0; h.job.len = n; h.job.first = 0; h.job.total = n; h.job.tile0 = 0; h.job.chunk0 = 0; h.job.blk0 = 0; h.job.halo = 0;
    h.prefix[0] = 0; h.prefix[1] = (int32_t)tiles;
    const size_t o_loc = (sizeof(Head) + 255) & ~size_t(255);
    const size_t o_abs = o_loc + sizeof(double) * (size_t)chunks;
    unsigned char* buf = nullptr;
    SC_CUDA(cudaMallocAsync((void**)&buf, o_abs + sizeof(double) * (size_t)chunks, st));
    SC_CUDA(cudaMemcpyAsync(buf, &h, sizeof(Head), cudaMemcpyHostToDevice, st));   // pageable source: staged before return
    const WavJob* dj = reinterpret_cast<const WavJob*>(buf);
    const int32_t* dp = reinterpret_cast<const int32_t*>(buf + offsetof(Head, prefix));
    double* loc = reinterpret_cast<double*>(buf + o_loc);
    double* cabs = reinterpret_cast<double*>(buf + o_abs);
    k_iir_local<TIN><<<(unsigned)tiles, kIirBlock, 0, st>>>(wav, dj, 1, dp, coeff, loc);
    SC_LAUNCHED();
    if (win == 0) {
        k_iir_carry<<<1, 32, 0, st>>>(dj, 1, coeff, loc);
        SC_LAUNCHED();
    }
    k_iir_apply<TIN><<<(unsigned)tiles, kIirBlock, 0, st>>>(wav, dj, 1, dp, coeff, iir_chunk_factor(coeff), win, loc, out, cabs);
    SC_LAUNCHED();
    SC_CUDA(cudaFreeAsync(buf, st));
    return SC_OK;
}
extern "C" int sc_inv_preemphasis(const float* wav, int64_t n, double coeff, double* out, void* stream) {
    if (!wav || !out || n < 0) return fail(SC_ERR_INVALID, "sc_inv_preemphasis: bad argument");
    if (n == 0) return SC_OK;
    return inv_preemph_one<float>(wav, n, coeff, out, (cudaStream_t)stream);
}
extern "C" int sc_inv_preemphasis_f64(const double* wav, int64_t n, double coeff, double* out, void* stream) {
    if (!wav || !out || n < 0) return fail(SC_ERR_INVALID, "sc_inv_preemphasis_f64: bad argument");
    if (n == 0) return SC_OK;
    return inv_preemph_one<double>(wav, n, coeff, out, (cudaStream_t)stream);
}

// ---- geometry of time-chunked runs: everything a caller needs to cut a long signal
static int64_t gl_out_per_tile(const sc_plan* pl) {
    return pl->fast ? kGlOut : gen_gl_out_per_tile(pl->prm.n_fft, pl->prm.hop_length);
}
static int64_t gcd64(int64_t a, int64_t b) { while (b) { const int64_t t = a % b; a = b; b = t; } return a; }
extern "C" int sc_chunk_geometry(const sc_plan* pl, int64_t* align_frames, int64_t* halo_samples, int64_t* halo_frames,
                                 int64_t* sum_block_samples) {
    if (!pl) return fail(SC_ERR_INVALID, "sc_chunk_geometry: null plan");
    const int64_t hop = pl->prm.hop_length, n_fft = pl->prm.n_fft;
    const int64_t tile_hops = gl_out_per_tile(pl) / hop;
    // cuts on the tile grid of the iteration kernel AND on the 256-sample grid of the de-emphasis scan
    const int64_t a = tile_hops * (kIirChunk / gcd64(kIirChunk, tile_hops * hop));
    if (align_frames) *align_frames = a;
    // one iteration couples a sample to audio within n_fft (frames whose centre is within n_fft/2, each n_fft long);
    // one more hop / frame keeps the pair partner of every needed frame exact (finding 3)
    const int64_t fr = (n_fft + hop - 1) / hop + 1;
    if (halo_frames) *halo_frames = fr;
    if (halo_samples) *halo_samples = fr * hop;
    if (sum_block_samples) *sum_block_samples = a * hop;
    return SC_OK;
}

// ------------------------------------------------------------------- de-emphasis + renormalisation
struct IirLayout {
    std::vector<WavJob> jobs;
    std::vector<int32_t> prefix, blk_prefix;
    int64_t loc_entries = 0;
    int blk_chunks = 1;
};
static int iir_layout(const sc_plan* pl, const std::vector<int64_t>& off, const std::vector<int64_t>& len,
                      const std::vector<int64_t>& first, const std::vector<int64_t>& total, int halo, IirLayout& L) {
    const int n = (int)off.size();
    int64_t a = 0, blk = 0;
    sc_chunk_geometry(pl, &a, nullptr, nullptr, &blk);
    L.blk_chunks = (int)(blk / kIirChunk);
    L.jobs.resize(n); L.prefix.assign(n + 1, 0); L.blk_prefix.assign(n + 1, 0);
    int64_t entries = 0;
    for (int u = 0; u < n; ++u) {
        WavJob& j = L.jobs[u];
        j.off = off[u]; j.len = len[u]; j.first = first[u]; j.total = total[u];
        j.tile0 = L.prefix[u]; j.chunk0 = (int32_t)entries; j.blk0 = L.blk_prefix[u]; j.halo = halo;
        const int64_t c = (len[u] + kIirChunk - 1) / kIirChunk;
        entries += halo + c;
        const int64_t t = L.prefix[u] + (c + kIirBlock - 1) / kIirBlock;
        const int64_t bb = L.blk_prefix[u] + (c + L.blk_chunks - 1) / L.blk_chunks;
        if (t > INT32_MAX || entries > INT32_MAX || bb > INT32_MAX) return fail(SC_ERR_INVALID, "signal too long");
        L.prefix[u + 1] = (int32_t)t; L.blk_prefix[u + 1] = (int32_t)bb;
    }
    L.loc_entries = entries;
    return 0;
}

extern "C" int sc_deemph_renorm_batch(sc_plan* pl, const float* wav, const int64_t* soff, const int64_t* slen_in,
                                      int32_t n, double coeff, double target, double* out, void* stream) {
    if (!pl || !wav || !soff || !out) return fail(SC_ERR_INVALID, "sc_deemph_renorm_batch: null argument");
    if (n <= 0) return SC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    SC_ENTER(pl, st, "sc_deemph_renorm_batch");
    std::vector<int64_t> off(soff, soff + n), len(n), first(n, 0);
    for (int u = 0; u < n; ++u) {
        len[u] = slen_in ? slen_in[u] : soff[u + 1] - soff[u];
        if (len[u] < 1) return fail(SC_ERR_INVALID, "sc_deemph_renorm_batch: empty signal");
    }
    IirLayout L;
    if (int rc = iir_layout(pl, off, len, first, len, 0, L)) return rc;
    Blob b;
    const size_t o_jobs = b.add(L.jobs), o_pre = b.add(L.prefix), o_blk = b.add(L.blk_prefix);
    if (int rc = upload_blob(pl, b, st)) return rc;
    const size_t w_abs = (sizeof(double) * (size_t)L.loc_entries + 255) & ~size_t(255);
    const size_t w_blk = (w_abs + sizeof(double) * (size_t)L.loc_entries + 255) & ~size_t(255);
    const size_t w_sc = (w_blk + sizeof(double) * (size_t)L.blk_prefix[n] + 255) & ~size_t(255);
    if (int rc = pl->work2.ensure(w_sc + sizeof(double) * n)) return rc;
    unsigned char* wb = static_cast<unsigned char*>(pl->work2.p);
    double* loc = reinterpret_cast<double*>(wb);
    double* cabs = reinterpret_cast<double*>(wb + w_abs);
    double* blks = reinterpret_cast<double*>(wb + w_blk);
    double* scale = reinterpret_cast<double*>(wb + w_sc);
    const WavJob* dj = at<WavJob>(pl, o_jobs);
    const int32_t* dp = at<int32_t>(pl, o_pre);
    const int32_t* db = at<int32_t>(pl, o_blk);
    const int win = iir_window(coeff);
    const int tiles = L.prefix[n];
    k_iir_local<float><<<tiles, kIirBlock, 0, st>>>(wav, dj, n, dp, coeff, loc);
    SC_LAUNCHED();
    if (win == 0) {
        k_iir_carry<<<(n + 31) / 32, 32, 0, st>>>(dj, n, coeff, loc);
        SC_LAUNCHED();
    }
    k_iir_apply<float><<<tiles, kIirBlock, 0, st>>>(wav, dj, n, dp, coeff, iir_chunk_factor(coeff), win, loc, out, cabs);
    SC_LAUNCHED();
    k_abs_blocks<<<(L.blk_prefix[n] + 127) / 128, 128, 0, st>>>(dj, n, db, L.blk_chunks, cabs, blks);
    SC_LAUNCHED();
    k_renorm_scale<<<(n + 127) / 128, 128, 0, st>>>(dj, n, db, blks, nullptr, 0, target, scale);
    SC_LAUNCHED();
    k_renorm<<<tiles, 256, 0, st>>>(dj, n, dp, scale, out);
    SC_LAUNCHED();
    return SC_OK;
}

// Time-chunked epilogue, three calls with two small exchanges between them (speechdsp.h):
//   sc_deemph_chunk_local  -> the caller sends its last `win` responses to the right neighbour
//   sc_deemph_chunk_apply  -> the caller all_gathers the block sums of |y|
//   sc_renorm_chunk
extern "C" int sc_deemph_chunk_window(double coeff) { return iir_window(coeff); }

static int chunk_job(sc_plan* pl, int64_t first, int64_t count, int64_t total, int halo, IirLayout& L, const char* who) {
    int64_t a = 0, blk = 0;
    sc_chunk_geometry(pl, &a, nullptr, nullptr, &blk);
    if (first < 0 || count < 1 || first + count > total) return fail(SC_ERR_INVALID, std::string(who) + ": bad sample range");
    if (first % blk != 0) return fail(SC_ERR_INVALID, std::string(who) + ": chunk start must be a multiple of sc_chunk_geometry's sum_block_samples");
    return iir_layout(pl, {0}, {count}, {first}, {total}, halo, L);
}

extern "C" int sc_deemph_chunk_local(sc_plan* pl, const float* wav, int64_t first, int64_t count, int64_t total,
                                     double coeff, double* loc_out, void* stream) {
    if (!pl || !wav || !loc_out) return fail(SC_ERR_INVALID, "sc_deemph_chunk_local: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    SC_ENTER(pl, st, "sc_deemph_chunk_local");
    IirLayout L;
    if (int rc = chunk_job(pl, first, count, total, 0, L, "sc_deemph_chunk_local")) return rc;
    Blob b;
    const size_t o_jobs = b.add(L.jobs), o_pre = b.add(L.prefix);
    if (int rc = upload_blob(pl, b, st)) return rc;
    k_iir_local<float><<<L.prefix[1], kIirBlock, 0, st>>>(wav, at<WavJob>(pl, o_jobs), 1, at<int32_t>(pl, o_pre), coeff, loc_out);
    SC_LAUNCHED();
    return SC_OK;
}

extern "C" int sc_deemph_chunk_apply(sc_plan* pl, const float* wav, int64_t first, int64_t count, int64_t total,
                                     double coeff, const double* loc_ext, int32_t n_halo, double* out,
                                     double* block_sums, void* stream) {
    if (!pl || !wav || !loc_ext || !out || !block_sums) return fail(SC_ERR_INVALID, "sc_deemph_chunk_apply: null argument");
    const int win = iir_window(coeff);
    if (win == 0) return fail(SC_ERR_UNSUPPORTED, "sc_deemph_chunk_apply: |coeff| is too close to 1 for the windowed carry; de-emphasise on one GPU");
    if (n_halo < 0 || (first > 0 && n_halo < win)) return fail(SC_ERR_INVALID, "sc_deemph_chunk_apply: need sc_deemph_chunk_window() halo entries from the left neighbour");
    cudaStream_t st = (cudaStream_t)stream;
    SC_ENTER(pl, st, "sc_deemph_chunk_apply");
    IirLayout L;
    if (int rc = chunk_job(pl, first, count, total, n_halo, L, "sc_deemph_chunk_apply")) return rc;
    Blob b;
    const size_t o_jobs = b.add(L.jobs), o_pre = b.add(L.prefix), o_blk = b.add(L.blk_prefix);
    if (int rc = upload_blob(pl, b, st)) return rc;
    if (int rc = pl->work2.ensure(sizeof(double) * (size_t)L.loc_entries)) return rc;
    double* cabs = static_cast<double*>(pl->work2.p);
    k_iir_apply<float><<<L.prefix[1], kIirBlock, 0, st>>>(wav, at<WavJob>(pl, o_jobs), 1, at<int32_t>(pl, o_pre), coeff,
                                                         iir_chunk_factor(coeff), win, loc_ext, out, cabs);
    SC_LAUNCHED();
    k_abs_blocks<<<(L.blk_prefix[1] + 127) / 128, 128, 0, st>>>(at<WavJob>(pl, o_jobs), 1, at<int32_t>(pl, o_blk), L.blk_chunks, cabs, block_sums);
    SC_LAUNCHED();
    return SC_OK;
}

extern "C" int sc_renorm_chunk(sc_plan* pl, double* out, int64_t count, const double* all_block_sums, int64_t n_blocks_total,
                               int64_t total, double target, void* stream) {
    if (!pl || !out || !all_block_sums) return fail(SC_ERR_INVALID, "sc_renorm_chunk: null argument");
    if (count < 1 || total < count || n_blocks_total < 1 || n_blocks_total > INT32_MAX) return fail(SC_ERR_INVALID, "sc_renorm_chunk: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    SC_ENTER(pl, st, "sc_renorm_chunk");
    IirLayout L;
    if (int rc = iir_layout(pl, {0}, {count}, {0}, {total}, 0, L)) return rc;
    Blob b;
    const size_t o_jobs = b.add(L.jobs), o_pre = b.add(L.prefix);
    if (int rc = upload_blob(pl, b, st)) return rc;
    if (int rc = pl->work2.ensure(sizeof(double))) return rc;
    double* scale = static_cast<double*>(pl->work2.p);
    k_renorm_scale<<<1, 128, 0, st>>>(at<WavJob>(pl, o_jobs), 1, nullptr, nullptr, all_block_sums, (int)n_blocks_total, target, scale);
    SC_LAUNCHED();
    k_renorm<<<L.prefix[1], 256, 0, st>>>(at<WavJob>(pl, o_jobs), 1, at<int32_t>(pl, o_pre), scale, out);
    SC_LAUNCHED();
    return SC_OK;
}

// --------------------------------------------------------------------------- power -> amp
struct P2aLayout {
    std::vector<P2aJob> jobs;
    std::vector<int32_t> prefix, blk_prefix;
    int block_rows = 1;
};
static int p2a_layout(const sc_plan* pl, const int64_t* foff, const int64_t* fcnt_in, int n, P2aLayout& L) {
    int64_t a = 0;
    sc_chunk_geometry(pl, &a, nullptr, nullptr, nullptr);
    L.block_rows = (int)a;
    L.jobs.resize(n); L.prefix.assign(n + 1, 0); L.blk_prefix.assign(n + 1, 0);
    for (int u = 0; u < n; ++u) {
        P2aJob& j = L.jobs[u];
        j.row0 = foff[u];
        j.rows = fcnt_in ? fcnt_in[u] : foff[u + 1] - foff[u];
        if (j.rows < 1) return fail(SC_ERR_INVALID, "power to amplitude: empty spectrogram");
        j.tile0 = L.prefix[u]; j.blk0 = L.blk_prefix[u];
        const int64_t t = L.prefix[u] + (j.rows * pl->n_bins + kP2aChunk - 1) / kP2aChunk;
        const int64_t bb = L.blk_prefix[u] + (j.rows + L.block_rows - 1) / L.block_rows;
        if (t > INT32_MAX || bb > INT32_MAX) return fail(SC_ERR_INVALID, "batch too large");
        L.prefix[u + 1] = (int32_t)t; L.blk_prefix[u + 1] = (int32_t)bb;
    }
    return 0;
}

extern "C" int sc_power_to_amp_batch(sc_plan* pl, const float* p, const int64_t* foff, const int64_t* fcnt_in,
                                     int32_t n, double norm, double realse, float* amp, void* stream) {
    if (!pl || !p || !foff || !amp) return fail(SC_ERR_INVALID, "sc_power_to_amp_batch: null argument");
    if (n <= 0) return SC_OK;
    if (norm == 0.0) return fail(SC_ERR_INVALID, "P_dB_norm_factor must be non-zero");
    cudaStream_t st = (cudaStream_t)stream;
    SC_ENTER(pl, st, "sc_power_to_amp_batch");
    P2aLayout L;
    if (int rc = p2a_layout(pl, foff, fcnt_in, n, L)) return rc;
    Blob b;
    const size_t o_j = b.add(L.jobs), o_p = b.add(L.prefix), o_b = b.add(L.blk_prefix);
    if (int rc = upload_blob(pl, b, st)) return rc;
    const size_t w_scale = (sizeof(double) * 2 * (size_t)L.blk_prefix[n] + 255) & ~size_t(255);
    if (int rc = pl->work2.ensure(w_scale + sizeof(float) * n)) return rc;
    double* partial = static_cast<double*>(pl->work2.p);
    float* scale = reinterpret_cast<float*>(static_cast<unsigned char*>(pl->work2.p) + w_scale);
    const int use_realse = realse != 1.0;
    if (use_realse) {
        k_p2a_partial<<<L.blk_prefix[n], 256, 0, st>>>(p, at<P2aJob>(pl, o_j), n, at<int32_t>(pl, o_b), pl->n_bins,
                                                       L.block_rows, (float)realse, partial);
        SC_LAUNCHED();
        k_p2a_scale<<<(n + 127) / 128, 128, 0, st>>>(n, at<int32_t>(pl, o_b), partial, scale);
        SC_LAUNCHED();
    }
    k_p2a_apply<<<L.prefix[n], 256, 0, st>>>(p, at<P2aJob>(pl, o_j), n, at<int32_t>(pl, o_p), pl->n_bins,
                                             (float)realse, use_realse, scale, (float)(1.0 / norm), amp);
    SC_LAUNCHED();
    return SC_OK;
}

// Time-chunked prologue: block partials of this rank's rows (rows [0, n_rows) of p_dev start at a multiple of
// sc_chunk_geometry's align_frames), then - after the caller gathered every rank's partials - the conversion.
extern "C" int sc_p2a_chunk_partial(sc_plan* pl, const float* p, int64_t n_rows, double realse, double* partial_out, void* stream) {
    if (!pl || !p || !partial_out || n_rows < 1) return fail(SC_ERR_INVALID, "sc_p2a_chunk_partial: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    SC_ENTER(pl, st, "sc_p2a_chunk_partial");
    P2aLayout L;
    const int64_t fo[2] = {0, n_rows};
    if (int rc = p2a_layout(pl, fo, nullptr, 1, L)) return rc;
    Blob b;
    const size_t o_j = b.add(L.jobs), o_b = b.add(L.blk_prefix);
    if (int rc = upload_blob(pl, b, st)) return rc;
    k_p2a_partial<<<L.blk_prefix[1], 256, 0, st>>>(p, at<P2aJob>(pl, o_j), 1, at<int32_t>(pl, o_b), pl->n_bins, L.block_rows,
                                                   (float)realse, partial_out);
    SC_LAUNCHED();
    return SC_OK;
}

extern "C" int sc_p2a_chunk_apply(sc_plan* pl, const float* p, int64_t n_rows, double norm, double realse,
                                  const double* all_partials, int64_t n_blocks_total, float* amp, void* stream) {
    if (!pl || !p || !amp || n_rows < 1) return fail(SC_ERR_INVALID, "sc_p2a_chunk_apply: bad argument");
    if (norm == 0.0) return fail(SC_ERR_INVALID, "P_dB_norm_factor must be non-zero");
    const int use_realse = realse != 1.0;
    if (use_realse && (!all_partials || n_blocks_total < 1 || n_blocks_total > INT32_MAX))
        return fail(SC_ERR_INVALID, "sc_p2a_chunk_apply: realse != 1 needs the gathered block partials");
    cudaStream_t st = (cudaStream_t)stream;
    SC_ENTER(pl, st, "sc_p2a_chunk_apply");
    P2aLayout L;
    const int64_t fo[2] = {0, n_rows};
    if (int rc = p2a_layout(pl, fo, nullptr, 1, L)) return rc;
    const int32_t all_prefix[2] = {0, (int32_t)(use_realse ? n_blocks_total : 0)};
    Blob b;
    const size_t o_j = b.add(L.jobs), o_p = b.add(L.prefix), o_all = b.add(all_prefix, 2);
    if (int rc = upload_blob(pl, b, st)) return rc;
    if (int rc = pl->work2.ensure(sizeof(float))) return rc;
    float* scale = static_cast<float*>(pl->work2.p);
    if (use_realse) {
        k_p2a_scale<<<1, 128, 0, st>>>(1, at<int32_t>(pl, o_all), all_partials, scale);
        SC_LAUNCHED();
    }
    k_p2a_apply<<<L.prefix[1], 256, 0, st>>>(p, at<P2aJob>(pl, o_j), 1, at<int32_t>(pl, o_p), pl->n_bins, (float)realse,
                                             use_realse, scale, (float)(1.0 / norm), amp);
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

// one launch of the one-tile-per-CTA kernels: initial inverse STFT (init) or a plain iteration
static int gl_launch(sc_plan* pl, bool init, const GlJob* jobs, int n, const int32_t* prefix, int n_tiles,
                     const float* amp, const float* phase0, const float* wav_in, float* wav_out, cudaStream_t st) {
    if (n_tiles == 0) return SC_OK;
    if (pl->fast) {
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
    const int half = pl->prm.n_fft / 2;
    const int64_t out_per_tile = gl_out_per_tile(pl);
    const int64_t p_first = ((out_first + half) / out_per_tile) * out_per_tile;
    const int64_t p_end = out_first + out_count + half;
    return out_count > 0 ? (p_end - p_first + out_per_tile - 1) / out_per_tile : 0;
}

extern "C" int sc_griffinlim_batch(sc_plan* pl, const float* amp, const float* phase0, const int64_t* foff,
                                   const int64_t* fcnt_in, int32_t n, int32_t n_iters, float* wav, const int64_t* soff,
                                   float* rms, void* stream) {
    if (!pl || !amp || !phase0 || !foff || !wav || !soff) return fail(SC_ERR_INVALID, "sc_griffinlim_batch: null argument");
    if (n_iters < 1) return fail(SC_ERR_INVALID, "sc_griffinlim_batch: n_iters must be >= 1 (the reference returns None for 0 iterations)");
    if (n <= 0) return SC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    SC_ENTER(pl, st, "sc_griffinlim_batch");
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
    const int n_sm = pl->n_sm;
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

// `n_steps` Griffin-Lim iterations on one rank's time chunk of a long signal, all queued on `stream` without a host
// synchronisation in between (speechdsp.h).  Step m reads wav_a (m even) or wav_b (m odd) and writes the other one
// over [own_first - (n_steps-1-m)*halo, own_first + own_count + (n_steps-1-m)*halo), clipped to the signal.
extern "C" int sc_griffinlim_chunk_run(sc_plan* pl, const float* amp, const float* phase0, int64_t first_frame,
                                       int64_t n_local, int64_t n_total, float* wav_a, float* wav_b, int64_t ext_first,
                                       int64_t ext_count, int64_t own_first, int64_t own_count, int32_t n_steps,
                                       int64_t halo_per_step, void* stream) {
    if (!pl || !amp || !wav_a || !wav_b) return fail(SC_ERR_INVALID, "sc_griffinlim_chunk_run: null argument");
    if (n_steps < 1) return fail(SC_ERR_INVALID, "sc_griffinlim_chunk_run: n_steps must be >= 1");
    if (n_total < 1 || (int64_t)pl->prm.hop_length * n_total + pl->prm.n_fft >= INT32_MAX || n_local < 0 || first_frame < 0 ||
        first_frame + n_local > n_total)
        return fail(SC_ERR_INVALID, "sc_griffinlim_chunk_run: bad frame range");
    const int64_t Lw = (int64_t)pl->prm.hop_length * (n_total - 1);
    if (own_first < 0 || own_count < 0 || own_first + own_count > Lw || ext_first < 0 || ext_count < 0 || ext_first + ext_count > Lw ||
        own_first < ext_first || own_first + own_count > ext_first + ext_count || halo_per_step < 0)
        return fail(SC_ERR_INVALID, "sc_griffinlim_chunk_run: bad sample ranges");
    if (own_count == 0) return SC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    SC_ENTER(pl, st, "sc_griffinlim_chunk_run");
    const int64_t ext_end = ext_first + ext_count;
    std::vector<GlJob> jobs(n_steps);
    std::vector<int32_t> prefix(2 * (size_t)n_steps);
    for (int m = 0; m < n_steps; ++m) {
        const int64_t grow = (int64_t)(n_steps - 1 - m) * halo_per_step;
        const int64_t lo = std::max(ext_first, own_first - grow), hi = std::min(ext_end, own_first + own_count + grow);
        GlJob& j = jobs[m];
        j.amp_row0 = 0; j.wav_in_off = 0; j.wav_in_first = ext_first; j.wav_in_count = ext_count;
        j.wav_out_off = lo - ext_first; j.out_first = lo; j.out_count = hi - lo;
        j.f_lo = (int32_t)first_frame; j.f_cnt = (int32_t)n_local; j.T = (int32_t)n_total; j.tile0 = 0;
        const int64_t tiles = gl_tiles_for(pl, lo, hi - lo);
        if (tiles > INT32_MAX) return fail(SC_ERR_INVALID, "chunk too large");
        prefix[2 * m] = 0; prefix[2 * m + 1] = (int32_t)tiles;
    }
    Blob b;
    const size_t o_j = b.add(jobs), o_p = b.add(prefix);
    if (int rc = upload_blob(pl, b, st)) return rc;
    const GlJob* dj = at<GlJob>(pl, o_j);
    const int32_t* dp = at<int32_t>(pl, o_p);
    for (int m = 0; m < n_steps; ++m) {
        const float* in = (m & 1) ? wav_b : wav_a;
        float* out = (m & 1) ? wav_a : wav_b;
        const int tiles = prefix[2 * m + 1];
        const bool init = m == 0 && phase0 != nullptr;
        if (pl->fast && !init) {
            // persistent iteration kernel (one job: no tile table needed)
            const int grid = tiles < 2 * pl->n_sm ? tiles : 2 * pl->n_sm;
            k_gl_iter_persist<<<grid, kFeThreads, sizeof(GlSmemP), st>>>(dj + m, nullptr, tiles, gl_tables(pl), amp, in, out);
            SC_LAUNCHED();
        } else if (int rc = gl_launch(pl, init, dj + m, 1, dp + 2 * m, tiles, amp, phase0, in, out, st)) return rc;
    }
    return SC_OK;
}

// single step with separate input / output buffers (the round-1 interface, kept for callers that own the exchange)
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
    SC_ENTER(pl, st, "sc_griffinlim_chunk_step");
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
        const int grid = tiles < 2 * pl->n_sm ? (int)tiles : 2 * pl->n_sm;
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
    const int n_ev = pl->pev_kind == 1 ? 4 : 3;
    SC_CUDA(cudaEventSynchronize(pl->pev[n_ev - 1]));
    for (int i = 0; i + 1 < n_ev; ++i) {
        float ms = 0.f;
        SC_CUDA(cudaEventElapsedTime(&ms, pl->pev[i], pl->pev[i + 1]));
        ms_out[i] = ms;
    }
    ms_out[3] = pl->pev_kind == 1 ? 1.0 : (double)pl->pev_iters;
    return SC_OK;
}

// Size every buffer of the plan for batches of up to max_utts utterances / signals, max_samples samples and
// max_frames feature rows in total, so that compute calls within these bounds do not allocate.
extern "C" int sc_plan_reserve(sc_plan* pl, int64_t max_samples, int64_t max_frames, int32_t max_utts) {
    if (!pl || max_samples < 0 || max_frames < 0 || max_utts < 0) return fail(SC_ERR_INVALID, "sc_plan_reserve: bad argument");
    PlanGuard guard_(pl);
    if (int rc = plan_enter(pl, guard_, pl->has_last ? pl->last_stream : nullptr, "sc_plan_reserve")) return rc;
    const int64_t n = max_utts, hop = pl->prm.hop_length;
    // rows as the packed layouts see them (every utterance start aligned to 4 rows) and samples of Griffin-Lim signals
    const int64_t rows = max_frames + 4 * n, gl_samples = hop * max_frames + 4 * n;
    const size_t heap = fe_heap_floats_bound(max_samples, n);
    // front-end
    if (int rc = pl->work.ensure(std::max(fe_work_bytes(pl, n, heap, rows) + sizeof(AbsRec) * heap,
                                          ((sizeof(float) * (size_t)gl_samples + 255) & ~size_t(255)) + sizeof(double) * (size_t)(gl_samples / 8192 + n + 1) + 8)))
        return rc;
    if (int rc = pl->fe_tab.ensure(fe_tab_bytes(rows / kWsFrames + n, rows / kB3Frames + n, (int64_t)heap))) return rc;
    if (int rc = pl->ds_fe.reserve(fe_desc_bytes(n))) return rc;
    // Griffin-Lim descriptors (jobs, prefixes, tile table), power -> amplitude partials, de-emphasis scan
    const int64_t gl_tiles = gl_samples / gl_out_per_tile(pl) + 2 * n + 2;
    const size_t ds_bytes = (size_t)n * (sizeof(GlJob) + sizeof(WavJob) + sizeof(P2aJob) + 64) + sizeof(int2) * (size_t)gl_tiles + 4096;
    if (int rc = pl->ds.reserve(std::max(ds_bytes, fe_desc_bytes(n)))) return rc;
    const int64_t chunks = gl_samples / kIirChunk + n + 1;
    int64_t a = 1;
    sc_chunk_geometry(pl, &a, nullptr, nullptr, nullptr);
    const size_t w2_iir = 3 * (((size_t)chunks * sizeof(double) + 255) & ~size_t(255)) + sizeof(double) * (size_t)n + 1024;
    const size_t w2_p2a = sizeof(double) * 2 * (size_t)(rows / a + n + 1) + sizeof(float) * (size_t)n + 1024;
    if (int rc = pl->work2.ensure(std::max(w2_iir, w2_p2a))) return rc;
    if (pl->pev.size() < 4) {
        for (int i = (int)pl->pev.size(); i < 4; ++i) {
            cudaEvent_t e = nullptr;
            SC_CUDA(cudaEventCreate(&e));
            pl->pev.push_back(e);
        }
    }
    return SC_OK;
}

// Device-side error flags raised by kernels since the last poll (bit 0: an utterance whose mean|y| is zero or whose
// gain is not finite - the reference's librosa.stft raises "not finite everywhere" for it).  Waits for `stream`.
extern "C" int sc_plan_poll_status(sc_plan* pl, int32_t* flags_out, void* stream) {
    if (!pl || !flags_out) return fail(SC_ERR_INVALID, "sc_plan_poll_status: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    SC_ENTER(pl, st, "sc_plan_poll_status");
    int32_t v = 0;
    SC_CUDA(cudaMemcpyAsync(&v, pl->status.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SC_CUDA(cudaStreamSynchronize(st));
    if (v != 0) SC_CUDA(cudaMemsetAsync(pl->status.p, 0, sizeof(int32_t), st));
    *flags_out = v;
    return SC_OK;
}

extern "C" int64_t sc_alloc_count(void) { return g_allocs.load(); }
extern "C" int64_t sc_launch_count(void) { return g_launches.load(); }
extern "C" void sc_launch_count_reset(void) { g_launches.store(0); }
extern "C" const char* sc_last_error(void) { return g_err.c_str(); }
extern "C" const char* sc_version(void) { return "speechdsp-b200 0.2 (sm_100a)"; }
