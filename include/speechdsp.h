/* speechdsp.h — C ABI of the B200-native audio-DSP hot path of socom20/speech-cloner.
 *
 * The reference has no FFI: its boundary is `from audio_lib import ...`.
 * Each entry point below is what a binding for that path would call; the reference
 * interface it replaces is cited as /root/reference file:line.  All pointers named *_dev
 * are CUDA device pointers owned by the caller, *_host are host pointers; `stream` is a
 * cudaStream_t passed as void*.  Every call is stream-ordered and returns 0 on success or a
 * negative sc_status; sc_last_error() gives the message.  No exceptions cross the ABI and
 * there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Threading / streams: a plan owns per-call state (descriptor staging, workspaces), so it is
 * SINGLE-THREADED - a second thread entering a call on the same plan gets SC_ERR_INVALID; create
 * one plan per thread.  Calls on one plan are ordered: when the stream changes between calls the
 * library makes the new stream wait for the work queued on the old one.  A plan belongs to the
 * device that was current in sc_plan_create; calling it with another device current is
 * SC_ERR_INVALID.  After sc_plan_reserve() compute calls within the reserved bounds do not
 * allocate (sc_alloc_count() counts the library's allocations; tests assert it stays put).
 * The only extension of sc_params over the reference's keyword arguments is fft_precision.
 *
 * Ragged batches: utterance u owns samples [sample_offsets[u], sample_offsets[u+1]) of the
 * packed waveform buffer and frames [frame_offsets[u], frame_offsets[u+1]) of the packed,
 * TIME-MAJOR feature buffers (row = frame, as audio_lib.py:207-211 returns them).  Gaps
 * between utterances are allowed (offsets[u+1]-offsets[u] may exceed the utterance) when the
 * explicit length arrays are given; see each call.
 */
#ifndef SPEECHDSP_H_
#define SPEECHDSP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sc_plan sc_plan;

typedef enum sc_status {
    SC_OK = 0,
    SC_ERR_INVALID = -1,     /* bad argument (ValueError on the Python side)           */
    SC_ERR_CUDA = -2,        /* CUDA runtime error, message has cudaGetErrorString     */
    SC_ERR_UNSUPPORTED = -3, /* parameter combination has no kernel                    */
    SC_ERR_NO_DEVICE = -4    /* no CUDA device: the library never computes on the CPU  */
} sc_status;

/* DSP parameters = the hp/*.json keys that change front-end results (the md5 key list of
 * TIMIT_reader.py:92-111) + calc_MFCC_input's keyword arguments (audio_lib.py:89-104). */
typedef struct sc_params {
    int32_t sample_rate;            /* sr                                   audio_lib.py:90  */
    int32_t n_fft;                  /* n_fft (None -> win_length upstream)  audio_lib.py:135 */
    int32_t win_length;             /* win_length                           audio_lib.py:93  */
    int32_t hop_length;             /* hop_length                           audio_lib.py:92  */
    int32_t n_mels;                 /* n_mels                               audio_lib.py:94  */
    int32_t n_mfcc;                 /* n_mfcc                               audio_lib.py:95  */
    int32_t mfcc_normalize_first;   /* mfcc_normaleze_first_mfcc            audio_lib.py:220 */
    int32_t calc_mfcc_derivative;   /* calc_mfcc_derivate                   audio_lib.py:226 */
    int32_t clip_output;            /* clip_output                          audio_lib.py:237 */
    int32_t fft_precision;          /* front-end FFT arithmetic: 0 = float64 like the reference's
                                       scipy FFT (default, meets 1e-5 abs), 1 = float32 (faster;
                                       bins 70-80 dB under the utterance max may be off by ~5e-5) */
    double pre_emphasis;            /* 0.0 disables the filter              audio_lib.py:129 */
    double mfcc_norm_factor;        /*                                      audio_lib.py:223 */
    double m_db_norm_factor;        /* 1.0 disables min-shift + scale       audio_lib.py:234 */
    double p_db_norm_factor;        /* 1.0 disables min-shift + scale       audio_lib.py:230 */
    double mean_abs_amp_norm;       /* 1.0 disables the gain                audio_lib.py:125 */
    const double* window_host;      /* win_length periodic window samples
                                       (scipy.signal.get_window(name, win, fftbins=True),
                                       audio_lib.py:145); NULL -> Hann                        */
} sc_params;

/* Build the constant tables (window, W400 twiddles, sparse Slaney mel filterbank, DCT-II
 * basis, window-sum-square) for one parameter set and upload them.  Replaces the per-call
 * librosa.filters.mel / librosa.filters.dct rebuilds of audio_lib.py:160-166, :176. */
int sc_plan_create(const sc_params* params, sc_plan** plan_out);
void sc_plan_destroy(sc_plan* plan);

/* Pre-size every buffer of the plan (workspaces, descriptor staging, pinned host slots) for batches of
 * up to max_utts utterances / spectrograms holding max_samples samples and max_frames frames in total.
 * Calls within these bounds then never call cudaMalloc / cudaFree / cudaMallocHost; larger calls still
 * work (they grow the buffers, which synchronises the device).  SURVEY.md section 8(b), last row. */
int sc_plan_reserve(sc_plan* plan, int64_t max_samples, int64_t max_frames, int32_t max_utts);

/* cudaMalloc + cudaMallocHost calls made by the library so far (all plans). */
int64_t sc_alloc_count(void);

/* Device-side error flags raised by kernels of this plan since the last poll; waits for `stream`.
 * bit 0: sc_frontend_batch met an utterance whose mean|y| is 0 or whose gain is not finite
 *        (audio_lib.py:126 divides by zero there and librosa.stft then raises "not finite everywhere"). */
int sc_plan_poll_status(sc_plan* plan, int32_t* flags_out, void* stream);

/* 1 if (n_fft, hop_length) take the hand-tuned FFT-400 kernels, 0 for the generic-size kernels. */
int sc_plan_is_fast_path(const sc_plan* plan);

/* Number of frames librosa.stft(center=True) yields for n_samples: 1 + n_samples / hop. */
int64_t sc_num_frames(const sc_plan* plan, int64_t n_samples);

/* calc_MFCC_input over a ragged batch (audio_lib.py:89-244; callers TIMIT_reader.py:175,
 * ARCTIC_reader.py:140, TARGET_spk_reader.py:156, test.py:474).
 *   wav_dev          packed float32 waveforms
 *   sample_offsets   n_utts+1 host int64; utterance u has sample_offsets[u+1]-sample_offsets[u]
 *                    samples unless sample_lengths_host != NULL (then that many, rest is padding)
 *   mfcc_dev         [frames][n_mfcc * (1 + calc_mfcc_derivative)] float32
 *   mel_dev          [frames][n_mels] float32
 *   pdb_dev          [frames][1 + n_fft/2] float32
 *   frame_offsets    n_utts+1 host int64 row offsets; rows per utterance = sc_num_frames()
 */
int sc_frontend_batch(sc_plan* plan, const float* wav_dev, const int64_t* sample_offsets_host,
                      const int64_t* sample_lengths_host, int32_t n_utts, float* mfcc_dev, float* mel_dev,
                      float* pdb_dev, const int64_t* frame_offsets_host, void* stream);

/* np.abs(y).mean() of every utterance as float32, bit-identical to NumPy's pairwise summation
 * (the gain of audio_lib.py:125-126 is mean_abs_amp_norm / this).  mean_out_dev: n_utts floats. */
int sc_mean_abs_batch(sc_plan* plan, const float* wav_dev, const int64_t* sample_offsets_host,
                      const int64_t* sample_lengths_host, int32_t n_utts, float* mean_out_dev, void* stream);

/* calc_PHN_target (audio_lib.py:51-85; callers TIMIT_reader.py:192, ARCTIC_reader.py:157,
 * TARGET_spk_reader.py:173) over a ragged batch: for every frame, the index (within its utterance's
 * interval list) of the phoneme interval that the reference's cursor + larger-overlap rule selects.
 *   phn_start_dev / phn_end_dev   packed int32 interval bounds in samples, [start, end)
 *   phn_offsets_host              n_utts+1: utterance u owns intervals [phn_offsets[u], phn_offsets[u+1]),
 *                                 at least one, ends in non-decreasing order (checked by the caller)
 *   sample_lengths_host           samples per utterance: frames = 1 + len / hop_length (:52)
 *   out_index_dev                 int32 per frame at row frame_offsets_host[u] + t
 * The caller maps indices to labels (phn_conv_d[phn_v[i][2]], :72-79). */
int sc_phn_target_batch(sc_plan* plan, const int32_t* phn_start_dev, const int32_t* phn_end_dev,
                        const int64_t* phn_offsets_host, const int64_t* sample_lengths_host, int32_t n_utts,
                        int32_t hop_length, int32_t win_length, int32_t* out_index_dev,
                        const int64_t* frame_offsets_host, void* stream);

/* Window samplers on a feature cache resident in device memory:
 * Sound_DS.spec_window_sampler (sound_ds.py:262-350) and TIMIT.window_sampler (TIMIT_reader.py:474-523) slice
 * `ds_h5py[group][i_sample][i_s:i_e]` out of every feature group and zero-pad short utterances
 * (sound_ds.py:246-259).  The random draws stay with the caller (NumPy's global generator, reference order);
 * this call copies the windows of one batch out of up to 4 packed buffers in one launch.
 *   src_dev_ptrs_host[a]   packed [n_rows_total][width[a]] buffer of 32-bit elements (float32 / int32), device
 *   dst_dev_ptrs_host[a]   dense [n_windows][n_timesteps][width[a]] output, device
 *   first_row_dev[w]       packed row of the first frame of window w (frame offset of the utterance + i_s)
 *   valid_rows_dev[w]      rows copied; rows valid_rows..n_timesteps-1 of the window are zero (padding);
 *                          windows reaching outside [0, n_rows_total) are cut, never read out of bounds */
int sc_window_gather(const void* const* src_dev_ptrs_host, void* const* dst_dev_ptrs_host,
                     const int64_t* width_host, int32_t n_arrays, int64_t n_rows_total,
                     const int64_t* first_row_dev, const int32_t* valid_rows_dev, int32_t n_windows,
                     int32_t n_timesteps, void* stream);

/* calc_preemphasis / calc_inv_preemphasis (audio_lib.py:12-28, :31-47): float32 in, float64 out
 * (scipy.signal.lfilter promotes), zero initial state, one signal of n samples. */
int sc_preemphasis(const float* wav_dev, int64_t n, double coeff, double* out_dev, void* stream);
int sc_inv_preemphasis(const float* wav_dev, int64_t n, double coeff, double* out_dev, void* stream);
/* the same for float64 input (lfilter keeps float64 input in float64) */
int sc_preemphasis_f64(const double* wav_dev, int64_t n, double coeff, double* out_dev, void* stream);
int sc_inv_preemphasis_f64(const double* wav_dev, int64_t n, double coeff, double* out_dev, void* stream);

/* Prologue of from_power_to_wav (audio_lib.py:290-298): P = max(0, P); optional `realse`
 * power law with mean preservation; A = sqrt(10^(0.1 * (P / p_db_norm_factor - 80))).
 * p_dev and amp_dev are time-major [frames][1 + n_fft/2]; may alias. */
int sc_power_to_amp_batch(sc_plan* plan, const float* p_dev, const int64_t* frame_offsets_host,
                          const int64_t* frame_counts_host, int32_t n_utts, double p_db_norm_factor,
                          double realse, float* amp_dev, void* stream);

/* griffin_lim_alg (audio_lib.py:249-274; only caller from_power_to_wav :299) over a ragged
 * batch, from an injected initial phase (the reference draws np.pi*np.random.rand on the host,
 * :255).  amp_dev / phase0_dev: time-major [frames][1 + n_fft/2] float32.  n_iters inverse
 * STFTs and n_iters-1 forward STFTs are run, exactly like the reference loop.
 *   wav_dev          packed float32 output, utterance u gets hop*(T_u-1) samples at sample_offsets[u]
 *   rms_delta_dev    NULL, or [n_utts][n_iters] float32 receiving sqrt(mean((last-wav)^2)) per
 *                    iteration (column 0 unused) — the value the reference prints when verbose (:262-264)
 */
int sc_griffinlim_batch(sc_plan* plan, const float* amp_dev, const float* phase0_dev,
                        const int64_t* frame_offsets_host, const int64_t* frame_counts_host, int32_t n_utts,
                        int32_t n_iters, float* wav_dev, const int64_t* sample_offsets_host,
                        float* rms_delta_dev, void* stream);

/* Epilogue of from_power_to_wav (audio_lib.py:301-306): optional de-emphasis IIR
 * y[n] = x[n] + c*y[n-1] in float64, then y * (mean_abs_amp_norm / mean|y|).  coeff == 0 skips
 * the filter.  wav_dev float32 in, out_dev float64 out, same ragged layout. */
int sc_deemph_renorm_batch(sc_plan* plan, const float* wav_dev, const int64_t* sample_offsets_host,
                           const int64_t* sample_lengths_host, int32_t n_utts, double coeff,
                           double mean_abs_amp_norm, double* out_dev, void* stream);

/* Layout helper for the reference's frequency-major arrays (librosa order, audio_lib.py:298
 * transposes P.T): src [rows][cols] -> dst [cols][rows]; src_is_f64 selects float64 input.
 * dst is always float32. */
int sc_transpose_to_f32(const void* src_dev, int32_t src_is_f64, int64_t rows, int64_t cols, float* dst_dev,
                        void* stream);

/* One Griffin-Lim projection step on time-chunked long-form audio:
 * same as one iteration of sc_griffinlim_batch for a single utterance, but the caller owns the
 * waveform state, so ranks can exchange halos between calls.
 *   first_frame / n_frames_total  position of this chunk inside the whole spectrogram
 *   amp_dev          [n_frames_local][bins] magnitudes of frames first_frame ..
 *   phase0_dev       NULL for a normal iteration; same shape as amp_dev for the initial inverse STFT
 *   wav_in_dev       whole-signal coordinates are implied: element i is sample wav_first + i
 *   wav_out_dev      receives samples [out_first, out_first + out_count)
 */
int sc_griffinlim_chunk_step(sc_plan* plan, const float* amp_dev, const float* phase0_dev, int64_t first_frame,
                             int64_t n_frames_local, int64_t n_frames_total, const float* wav_in_dev,
                             int64_t wav_first, int64_t wav_count, float* wav_out_dev, int64_t out_first,
                             int64_t out_count, void* stream);

/* ---- time-chunked long-form path (BASELINE.json configs[3]) ------------------
 * One rank owns the samples [own_first, own_first + own_count) of a signal of n_frames_total frames
 * (hop * (n_frames_total - 1) samples).  sc_chunk_geometry gives the constants a caller cuts with:
 *   align_frames       cut points must be multiples of this many hops (tile grid of the iteration
 *                      kernel and 256-sample grid of the de-emphasis scan)
 *   halo_samples/_frames  reach of ONE iteration on each side (samples of waveform, rows of magnitude)
 *   sum_block_samples  block size of the canonical |y| / power sums (= hop * align_frames)
 * Everything below is computed on the whole-signal grid, so the assembled result is bit-identical
 * to the single-GPU one for any number of ranks. */
int sc_chunk_geometry(const sc_plan* plan, int64_t* align_frames, int64_t* halo_samples, int64_t* halo_frames,
                      int64_t* sum_block_samples);

/* n_steps Griffin-Lim iterations (audio_lib.py:259-270) on this rank's chunk, queued back to back on
 * `stream`.  wav_a_dev / wav_b_dev both cover samples [ext_first, ext_first + ext_count), the chunk
 * plus n_steps * halo_per_step samples each side (less at the signal ends); amp_dev (and phase0_dev)
 * hold rows first_frame .. first_frame + n_frames_local.  Step m reads a (m even) or b (m odd) and
 * writes the other over the chunk widened by (n_steps - 1 - m) * halo_per_step: the caller exchanges
 * halos once per call instead of once per iteration (communication-avoiding form of the per-iteration
 * exchange in SURVEY.md section 8(e)).  phase0_dev != NULL makes step 0 the initial inverse STFT (:256-260),
 * which does not read a waveform.  The result is in b when n_steps is odd, in a when it is even. */
int sc_griffinlim_chunk_run(sc_plan* plan, const float* amp_dev, const float* phase0_dev, int64_t first_frame,
                            int64_t n_frames_local, int64_t n_frames_total, float* wav_a_dev, float* wav_b_dev,
                            int64_t ext_first, int64_t ext_count, int64_t own_first, int64_t own_count,
                            int32_t n_steps, int64_t halo_per_step, void* stream);

/* Prologue (audio_lib.py:290-298) on a chunk.  With realse != 1 the two means of :293/:296 span the whole
 * signal: sc_p2a_chunk_partial writes 2 float64 per block of align_frames rows (sum P, sum P^realse) of
 * this rank's rows, the caller all_gathers them in rank order and passes all n_blocks_total pairs on. */
int sc_p2a_chunk_partial(sc_plan* plan, const float* p_dev, int64_t n_rows, double realse, double* partial_out_dev,
                         void* stream);
int sc_p2a_chunk_apply(sc_plan* plan, const float* p_dev, int64_t n_rows, double p_db_norm_factor, double realse,
                       const double* all_partials_dev, int64_t n_blocks_total, float* amp_dev, void* stream);

/* Epilogue (audio_lib.py:301-306) on a chunk [first, first + count) of a signal of `total` samples.
 *   sc_deemph_chunk_window  entries of loc a rank needs from its left neighbour (0: coefficient too close
 *                           to 1 for the windowed carry - de-emphasise on one GPU)
 *   sc_deemph_chunk_local   loc_out_dev[k] = zero-state response at the end of 256-sample chunk k
 *   sc_deemph_chunk_apply   loc_ext_dev = n_halo entries from the left neighbour (its last ones; zeros
 *                           for the first rank) followed by this rank's loc; writes float64 samples and
 *                           one float64 sum of |y| per block of sum_block_samples
 *   sc_renorm_chunk         y *= target / (sum of ALL ranks' block sums in order / total) */
int sc_deemph_chunk_window(double coeff);
int sc_deemph_chunk_local(sc_plan* plan, const float* wav_dev, int64_t first, int64_t count, int64_t total,
                          double coeff, double* loc_out_dev, void* stream);
int sc_deemph_chunk_apply(sc_plan* plan, const float* wav_dev, int64_t first, int64_t count, int64_t total,
                          double coeff, const double* loc_ext_dev, int32_t n_halo, double* out_dev,
                          double* block_sums_dev, void* stream);
int sc_renorm_chunk(sc_plan* plan, double* out_dev, int64_t count, const double* all_block_sums_dev,
                    int64_t n_blocks_total, int64_t total, double mean_abs_amp_norm, void* stream);

/* Per-kernel device timing for the roofline report (bench.py).  When enabled, the next
 * sc_frontend_batch / sc_griffinlim_batch records CUDA events between its kernels on the caller's
 * stream; sc_profile_read waits for them and returns milliseconds:
 *   after sc_frontend_batch:   ms[0] gain (|y| mean), ms[1] pass A (STFT..mel), ms[2] pass B, ms[3] = 1
 *   after sc_griffinlim_batch: ms[0] initial inverse STFT, ms[1] the n_iters-1 iterations, ms[2] 0, ms[3] = n_iters */
int sc_profile_enable(sc_plan* plan, int32_t on);
int sc_profile_read(sc_plan* plan, double* ms_out4);

/* Kernel launches issued by this library since the last reset (bench.py's gpu_launches). */
int64_t sc_launch_count(void);
void sc_launch_count_reset(void);

/* Message of the last failing call on this thread ("" if none). */
const char* sc_last_error(void);

/* Library version string. */
const char* sc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SPEECHDSP_H_ */
