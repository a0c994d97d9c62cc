"""CPU oracle for the speech-cloner audio-DSP hot path.  TEST INFRASTRUCTURE ONLY.

This module is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``speech_cloner_b200``) never imports anything
from ``oracle/`` and fails loudly when its CUDA library is missing.

PARITY PINNED TO THE REFERENCE SOURCE FOR ITS OWN CODE, UNPINNED FOR LIBROSA'S.  The reference
(``/root/reference/audio_lib.py``) delegates its transforms to librosa (un-vendored, un-pinned,
0.6.x by API usage: ``librosa.filters.dct`` at audio_lib.py:176, ``librosa.output.write_wav`` at
test.py:177) and scipy.signal; librosa cannot be imported or installed in this image and the
reference ships no tests / golden vectors.  So this file restates

  * ``audio_lib.py:12-47``   pre-/de-emphasis                (scipy.signal.lfilter)
  * ``audio_lib.py:51-85``   phoneme frame labels            (integer logic)
  * ``audio_lib.py:89-244``  ``calc_MFCC_input``             (librosa stft / power_to_db /
                             filters.mel / amplitude_to_db / filters.dct)
  * ``audio_lib.py:249-274`` ``griffin_lim_alg``             (librosa istft / stft / magphase)
  * ``audio_lib.py:278-308`` ``from_power_to_wav``           (librosa db_to_power)

with the published librosa-0.6 algorithms, written with explicit dtypes so that NumPy 2
scalar-promotion rules cannot change a result.  Two kinds of pins:

  1. the reference file ITSELF, imported unmodified from /root/reference with a librosa shim in
     sys.modules (``tests/golden/librosa_shim.py``), is run on seeded inputs by
     ``tests/golden/make_reference_vectors.py``; its outputs are frozen in
     ``tests/golden/reference_run_vectors.npz`` and this oracle reproduces them bit for bit
     (``tests/test_reference_run.py``).  That pins everything audio_lib.py does in its own code:
     gain, emphasis filters, dtype chain, normalisation, deltas, clipping, the Griffin-Lim loop with
     its ``np.random.rand`` phase, ``realse``, the label loop.
  2. the librosa primitives underneath (stft, istft, filters.mel, filters.dct, the dB helpers) have
     no reference-held check; they are pinned piecewise against independent implementations in
     ``tests/test_oracle_pins.py`` (torch.stft / istft, transformers.audio_utils, scipy.fft.dct,
     torchaudio) and by the frozen known-answer vectors in ``tests/golden/oracle_vectors.npz``.

dtype chain restated from the reference era (NumPy 1.x + librosa 0.6 + scipy.fftpack):
  gain            float32 array * float32 scalar                      audio_lib.py:126
  pre-emphasis    lfilter promotes to float64                         audio_lib.py:27
  STFT            float64 window * frames -> float64 FFT -> complex64 audio_lib.py:141
  P, P_dB         float32                                             audio_lib.py:150-157
  mel, M_dB, MFCC float64 (float64 filterbank @ float32 power)        audio_lib.py:160-179
  outputs         cast to float32                                     audio_lib.py:244
  Griffin-Lim     initial state complex128, later states complex64, iSTFT buffer float32,
                  window-sum-square float32                           audio_lib.py:255-270
"""
from __future__ import annotations

import numpy as np
import scipy.fftpack as _fftpack
from scipy import signal as _signal

__all__ = [
    "calc_preemphasis", "calc_inv_preemphasis", "calc_PHN_target", "calc_MFCC_input",
    "griffin_lim_alg", "from_power_to_wav",
    "stft", "istft", "mel_filterbank", "dct_basis", "power_to_db", "amplitude_to_db",
    "window_sumsquare", "padded_window", "reflect_index",
]

_F32_TINY = np.finfo(np.float32).tiny


# --------------------------------------------------------------------------- helpers
def padded_window(window, win_length: int, n_fft: int) -> np.ndarray:
    """Periodic window of ``win_length`` centred inside ``n_fft`` zeros (float64).

    librosa 0.6 ``filters.get_window(window, win_length, fftbins=True)`` followed by
    ``util.pad_center(.., n_fft)``; call sites audio_lib.py:141-147, :260, :267.
    """
    if callable(window):
        w = np.asarray(window(win_length), dtype=np.float64)
    elif isinstance(window, (str, tuple)) or np.isscalar(window):
        w = _signal.get_window(window, win_length, fftbins=True).astype(np.float64)
    else:
        w = np.asarray(window, dtype=np.float64)
        if w.shape != (win_length,):
            raise ValueError("window array must have length win_length")
    if win_length > n_fft:
        raise ValueError("win_length must be <= n_fft")
    lpad = (n_fft - win_length) // 2
    out = np.zeros(n_fft, dtype=np.float64)
    out[lpad:lpad + win_length] = w
    return out


def reflect_index(idx: np.ndarray, n: int) -> np.ndarray:
    """Index map of ``np.pad(.., mode='reflect')`` for arbitrary (also > n-1) pad widths."""
    idx = np.asarray(idx, dtype=np.int64)
    if n == 1:
        return np.zeros_like(idx)
    period = 2 * (n - 1)
    m = np.mod(idx, period)
    return np.where(m >= n, period - m, m)


def _check_audio(y) -> np.ndarray:
    """librosa ``util.valid_audio`` as reached through stft (SURVEY §8(b) 'Errors')."""
    if not isinstance(y, np.ndarray):
        raise ValueError("audio must be a numpy.ndarray")
    if not np.issubdtype(y.dtype, np.floating):
        raise ValueError("audio must be floating point")
    if y.ndim != 1:
        raise ValueError("audio must be one-dimensional (mono)")
    if y.shape[0] == 0:
        raise ValueError("audio is empty")
    if not np.isfinite(y).all():
        raise ValueError("audio is not finite everywhere")
    return y


def stft(y: np.ndarray, n_fft: int, hop_length: int, win_length=None, window="hann") -> np.ndarray:
    """librosa-0.6 ``core.stft(center=True, pad_mode='reflect', dtype=complex64)``.

    Returns ``(1 + n_fft//2, 1 + len(y)//hop)`` complex64, frequency-major.
    """
    if win_length is None:
        win_length = n_fft
    _check_audio(y)
    w = padded_window(window, win_length, n_fft)
    n = y.shape[0]
    n_frames = 1 + n // hop_length
    # frame t covers padded[t*hop : t*hop + n_fft], padded = reflect-pad by n_fft//2
    pos = (np.arange(n_frames, dtype=np.int64)[:, None] * hop_length
           + np.arange(n_fft, dtype=np.int64)[None, :] - n_fft // 2)
    frames = y[reflect_index(pos, n)]                       # (T, n_fft), dtype of y
    spec = _fftpack.fft(w[None, :] * frames, axis=1)        # float64 window promotes -> f64 FFT
    return np.ascontiguousarray(spec[:, : 1 + n_fft // 2].T).astype(np.complex64)


def window_sumsquare(window, n_frames: int, hop_length: int, win_length: int, n_fft: int) -> np.ndarray:
    """librosa-0.6 ``filters.window_sumsquare(.., dtype=float32, norm=None)``.

    A float32 accumulator receives float64 squared-window terms frame by frame in
    ascending frame order (``x[a:b] += win_sq[..]``), so every add rounds to float32.
    """
    n = n_fft + hop_length * (n_frames - 1)
    wsq = padded_window(window, win_length, n_fft) ** 2
    return _ordered_ola_f32(np.broadcast_to(wsq, (n_frames, n_fft)), hop_length, n)


def _ordered_ola_f32(frames_f64: np.ndarray, hop: int, n: int) -> np.ndarray:
    """Overlap-add of float64 frames into a float32 buffer, ascending frame order per sample.

    Equivalent to the reference-era loop ``y[s:s+n_fft] = y[s:s+n_fft] + ytmp`` (each add is
    float64, the store rounds to float32) but vectorised over samples: pass ``j`` adds, to
    every sample, the contribution of its ``j``-th covering frame.
    """
    n_frames, n_fft = frames_f64.shape
    s = np.arange(n, dtype=np.int64)
    first = np.maximum(0, -((n_fft - 1 - s) // hop))        # ceil((s-n_fft+1)/hop) clipped to 0
    last = np.minimum(n_frames - 1, s // hop)
    y = np.zeros(n, dtype=np.float32)
    depth = int((last - first).max()) + 1 if n else 0
    for j in range(depth):
        i = first + j
        ok = i <= last
        ic = np.where(ok, i, 0)
        term = np.where(ok, frames_f64[ic, np.where(ok, s - ic * hop, 0)], 0.0)
        y = np.where(ok, (y.astype(np.float64) + term).astype(np.float32), y)
    return y


def istft(spec: np.ndarray, hop_length: int, win_length=None, window="hann") -> np.ndarray:
    """librosa-0.6 ``core.istft(center=True, dtype=float32, length=None)``.

    ``spec`` is ``(1 + n_fft//2, T)``; complex128 input runs a float64 inverse FFT,
    complex64 input a float32 one (scipy.fftpack behaviour the reference relied on).
    """
    n_bins, n_frames = spec.shape
    n_fft = 2 * (n_bins - 1)
    if win_length is None:
        win_length = n_fft
    w = padded_window(window, win_length, n_fft)
    full = np.concatenate([spec, np.conj(spec[-2:0:-1])], axis=0)      # Hermitian extension
    frames = w[None, :] * _fftpack.ifft(full.T, axis=1).real           # (T, n_fft) float64
    n = n_fft + hop_length * (n_frames - 1)
    y = _ordered_ola_f32(frames, hop_length, n)
    wss = window_sumsquare(window, n_frames, hop_length, win_length, n_fft)
    nz = wss > _F32_TINY
    y[nz] = y[nz] / wss[nz]
    return y[n_fft // 2: n - n_fft // 2]


def power_to_db(S: np.ndarray, amin: float = 1e-10, top_db: float = 80.0) -> np.ndarray:
    """librosa-0.6 ``core.power_to_db(S, ref=1.0)``; result keeps the dtype of ``S``."""
    S = np.asarray(S)
    dt = S.dtype.type
    db = dt(10.0) * np.log10(np.maximum(dt(amin), S))
    if top_db is not None:
        db = np.maximum(db, dt(db.max() - dt(top_db)))
    return db


def amplitude_to_db(S: np.ndarray, top_db: float = 80.0) -> np.ndarray:
    """librosa-0.6 ``core.amplitude_to_db(S, ref=1.0, amin=1e-5)`` = power_to_db(S**2, amin=amin**2)."""
    mag = np.abs(np.asarray(S))
    return power_to_db(np.square(mag), amin=1e-5 ** 2, top_db=top_db)


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    lin = f / f_sp
    with np.errstate(divide="ignore"):
        log = min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep
    return np.where(f >= min_log_hz, log, lin)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft: int, n_mels: int, fmin: float = 0.0, fmax=None) -> np.ndarray:
    """librosa-0.6 ``filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm=1)`` (float64)."""
    if fmax is None:
        fmax = float(sr) / 2
    n_bins = 1 + n_fft // 2
    fft_f = np.linspace(0.0, float(sr) / 2, n_bins, endpoint=True)
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    edges = _mel_to_hz_slaney(mel_pts)
    width = np.diff(edges)
    ramps = edges[:, None] - fft_f[None, :]
    rise = -ramps[:-2] / width[:-1, None]
    fall = ramps[2:] / width[1:, None]
    tri = np.maximum(0.0, np.minimum(rise, fall))
    tri *= (2.0 / (edges[2:] - edges[:-2]))[:, None]       # Slaney area normalisation
    return tri


def dct_basis(n_out: int, n_in: int) -> np.ndarray:
    """librosa-0.6 ``filters.dct(n_out, n_in)``: orthonormal DCT-II rows 0..n_out-1 (float64)."""
    k = np.arange(1, 2 * n_in, 2, dtype=np.float64) * np.pi / (2.0 * n_in)
    basis = np.cos(np.arange(n_out, dtype=np.float64)[:, None] * k[None, :]) * np.sqrt(2.0 / n_in)
    basis[0, :] = 1.0 / np.sqrt(n_in)
    return basis


# --------------------------------------------------------------------- reference API
def calc_preemphasis(wav, coeff=0.97):
    """audio_lib.py:12-28 — ``lfilter([1, -coeff], [1], wav)``; zero initial state, float64 out."""
    x = np.asarray(wav, dtype=np.float64)
    y = x.copy()
    y[1:] -= np.float64(coeff) * x[:-1]
    return y


def calc_inv_preemphasis(preem_wav, coeff=0.97):
    """audio_lib.py:31-47 — ``lfilter([1], [1, -coeff], x)``: y[n] = x[n] + coeff*y[n-1], float64."""
    return _signal.lfilter([1.0], [1.0, -float(coeff)], np.asarray(preem_wav, dtype=np.float64))


def calc_PHN_target(y, phn_v, phn_conv_d, hop_length=40, win_length=400):
    """audio_lib.py:51-85 — per-frame phoneme label by larger overlap with the analysis window.

    ``phn_v`` is a list of ``(start_sample, end_sample, symbol)``; ``phn_conv_d`` maps a
    symbol to its label (int or one-hot vector).  Frame ``t`` looks at
    ``[t*hop - win//2, t*hop + win - win//2)``; the cursor advances while the current
    interval ends at or before the window start; the current and next intervals compete,
    ties go to the current one.
    """
    n_frames = int(y.shape[0] / hop_length) + 1
    half = win_length // 2
    labels = []
    cur = 0
    last = len(phn_v) - 1
    for t in range(n_frames):
        lo = t * hop_length - half
        hi = lo + win_length
        while phn_v[cur][1] <= lo and cur < last:
            cur += 1
        pick = cur
        if cur < last:
            ov_cur = min(phn_v[cur][1], hi) - max(phn_v[cur][0], lo)
            ov_nxt = min(phn_v[cur + 1][1], hi) - max(phn_v[cur + 1][0], lo)
            if ov_cur < ov_nxt:
                pick = cur + 1
        labels.append(phn_conv_d[phn_v[pick][2]])
    return np.array(labels, dtype=np.int32)


def calc_MFCC_input(y,
                    sr=16000,
                    pre_emphasis=0.97,
                    hop_length=40,
                    win_length=400,
                    n_mels=128,
                    n_mfcc=40,
                    n_fft=None,
                    window='hann',
                    mfcc_normaleze_first_mfcc=True,
                    mfcc_norm_factor=0.01,
                    calc_mfcc_derivate=False,
                    M_dB_norm_factor=0.01,
                    P_dB_norm_factor=0.01,
                    mean_abs_amp_norm=0.003,
                    clip_output=True):
    """audio_lib.py:89-244.  Returns (MFCC (T, n_mfcc[*2]), M_dB (T, n_mels), P_dB (T, 1+n_fft//2)) float32."""
    _check_audio(y)
    if mean_abs_amp_norm != 1.0:                                          # :125-126
        # NumPy-1.x era promotion: python float / float32 scalar -> float64 scalar, which is
        # then cast to the array dtype before the multiply.
        g = np.float64(mean_abs_amp_norm) / np.float64(np.abs(y).mean())
        y = y * y.dtype.type(g)
    y_pe = calc_preemphasis(y, pre_emphasis) if pre_emphasis != 0.0 else y  # :129-133
    if n_fft is None:                                                     # :135-136
        n_fft = win_length

    F = stft(y_pe, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window)  # :141-147
    mag = np.abs(F)                                                       # :150   float32
    P = mag * mag                                                         # :155   float32
    P_dB = power_to_db(P)                                                 # :157   float32

    M = mel_filterbank(sr, n_fft, n_mels)                                 # :160-166 float64
    M_spec = M @ P.astype(np.float64)                                     # :169
    M_dB = amplitude_to_db(M_spec)                                        # :172   float64
    MFCC = dct_basis(n_mfcc, n_mels) @ M_dB                               # :176-179

    MFCC = np.array(MFCC.T)                                               # :207-211 time-major
    M_dB = M_dB.T
    P_dB = P_dB.T

    if mfcc_normaleze_first_mfcc:                                         # :220-221
        MFCC[:, 0] -= MFCC[0, 0]
    if mfcc_norm_factor != 1.0:                                           # :223-224
        MFCC = mfcc_norm_factor * MFCC
    if calc_mfcc_derivate:                                                # :226-228
        if MFCC.shape[0] < 2:
            raise ValueError("calc_mfcc_derivate needs at least 2 frames")
        d = np.zeros_like(MFCC)
        d[1:-1] = 2.0 * (MFCC[2:] - MFCC[:-2])
        MFCC = np.concatenate([MFCC, d], axis=1)
    if P_dB_norm_factor != 1.0:                                           # :230-231 float32
        P_dB = np.float32(P_dB_norm_factor) * (P_dB - P_dB.min())
    if M_dB_norm_factor != 1.0:                                           # :234-235 float64
        M_dB = M_dB_norm_factor * (M_dB - M_dB.min())
    if clip_output:                                                       # :237-240
        MFCC = np.clip(MFCC, -1.0, 1.0)
        P_dB = np.clip(P_dB, -1.0, 1.0)
        M_dB = np.clip(M_dB, -1.0, 1.0)
    return (np.ascontiguousarray(MFCC, dtype=np.float32),
            np.ascontiguousarray(M_dB, dtype=np.float32),
            np.ascontiguousarray(P_dB, dtype=np.float32))


def griffin_lim_alg(stft_amp, win_length, hop_length, num_iters=300, n_fft=None, verbose=True,
                    phase0=None, rms_log=None):
    """audio_lib.py:249-274.  ``stft_amp`` is (1+n_fft//2, T) frequency-major.

    ``phase0`` (extra keyword) injects the initial phase; ``None`` draws
    ``np.pi * np.random.rand(*stft_amp.shape)`` from the global state exactly as :255 does.
    iSTFT/STFT use librosa's default Hann window regardless of the front-end window.
    ``rms_log`` (list) receives the per-iteration RMS deltas the reference prints (:262-264).
    """
    if n_fft is None:
        n_fft = win_length
    stft_amp = np.asarray(stft_amp)
    if phase0 is None:
        phase0 = np.pi * np.random.rand(*stft_amp.shape)
    S = stft_amp * np.exp(1.0j * np.asarray(phase0, dtype=np.float64))     # :256 complex128
    wav = last = None
    for i in range(num_iters):
        wav = istft(S, hop_length=hop_length, win_length=win_length)       # :260 float32
        if last is not None and (verbose or rms_log is not None):
            d = np.sqrt(np.mean(np.square(last - wav)))
            if rms_log is not None:
                rms_log.append(float(d))
            if verbose:
                print(' i={}  mrse_delta = {}'.format(i, d))
        if i != num_iters - 1:
            X = stft(wav, n_fft=n_fft, hop_length=hop_length, win_length=win_length)   # :267 complex64
            unit = np.exp(np.complex64(1.0j) * np.angle(X))                # :268 magphase, complex64
            ang = np.angle(unit)                                           # :269 float32
            S = stft_amp * np.exp(np.complex64(1.0j) * ang)                # :270 complex64 for f32 amp
        last = wav
    return wav


def from_power_to_wav(P,
                      P_dB_norm_factor=0.01,
                      pre_emphasis=0.97,
                      hop_length=40,
                      win_length=800,
                      mean_abs_amp_norm=0.01,
                      n_iter=200,
                      n_fft=None,
                      realse=1.0,
                      verbose=True,
                      phase0=None):
    """audio_lib.py:278-308.  ``P`` is the time-major (T, 1+n_fft//2) normalised power-dB map."""
    P = np.maximum(np.float32(0.0), np.asarray(P, dtype=np.float32))       # :290
    if realse != 1.0:                                                      # :292-296
        p_mean = P.mean()
        P = P ** np.float32(realse)
        P = (p_mean / P.mean()) * P
    dB = P.T / np.float32(P_dB_norm_factor) - np.float32(80.0)             # :298
    amp = np.sqrt(np.power(np.float32(10.0), np.float32(0.1) * dB))
    y = griffin_lim_alg(amp, win_length, hop_length, num_iters=n_iter, n_fft=n_fft,
                        verbose=verbose, phase0=phase0)                    # :299
    if pre_emphasis != 0:                                                  # :301-304
        y = calc_inv_preemphasis(y, pre_emphasis)
    return y * (mean_abs_amp_norm / np.abs(y).mean())                      # :306


def compound(y0, y1):
    """test.py:46-84 — stitch window predictions: ``y0`` (N, T, X) on the window grid, ``y1`` (N-1, T, X) on the
    half-offset grid; the middle halves alternate between the two, the first / last window keep their outer 3/4.
    Literal transcription (test infrastructure for speech_cloner_b200.conversion)."""
    n_quarter = y0.shape[1] // 4
    i_0, i_1 = 1, 0
    y_v = [y0[0, :-n_quarter, :]]
    while True:
        do_break = True
        if i_1 < y1.shape[0]:
            y_v.append(y1[i_1, n_quarter:-n_quarter, :])
            i_1 += 1
            do_break = False
        if i_0 < y0.shape[0] - 1:
            y_v.append(y0[i_0, n_quarter:-n_quarter, :])
            i_0 += 1
            do_break = False
        if do_break:
            break
    y_v.append(y0[-1, n_quarter:, :])
    return np.concatenate(y_v, axis=0)


def normalize_wav(y):
    """librosa.output.write_wav(..., norm=True) (test.py:177-179) = librosa.util.normalize(y, norm=inf):
    divide by max|y| unless it is below the dtype's tiny."""
    y = np.asarray(y)
    mag = np.max(np.abs(y)) if y.size else 0.0
    tiny = np.finfo(y.dtype if np.issubdtype(y.dtype, np.floating) else np.float32).tiny
    return y if mag < tiny else y / mag


def window_batches(mfcc, mel, stft, cfg_d, t_s=5, t_e=60):
    """Literal transcription of the batching in conversion2 (test.py:92-128): returns what is fed to decoder.predict."""
    hop = cfg_d['hop_length']
    n_times = cfg_d['n_timesteps']
    if mfcc.shape[0] % n_times != 0:
        pad_len = (n_times) - (mfcc.shape[0] % n_times)
        pad_mfcc = np.zeros((pad_len, mfcc.shape[1]))
        mfcc = np.concatenate([mfcc, pad_mfcc], axis=0)
        pad_mel = np.zeros((pad_len, mel.shape[1]))
        mel = np.concatenate([mel, pad_mel], axis=0)
        pad_stft = np.zeros((pad_len, stft.shape[1]))
        stft = np.concatenate([stft, pad_stft], axis=0)
    n_hop_s = t_s * cfg_d['sample_rate'] // hop
    n_hop_e = t_e * cfg_d['sample_rate'] // hop
    n_hop_e = min(n_hop_e, mfcc.shape[0])
    n_delta = n_times * ((n_hop_e - n_hop_s) // n_times)
    n_s = n_hop_s
    n_e = n_hop_s + n_delta
    if n_e <= n_s:
        raise Exception(' - ERROR, translate: n_e <= n_s.')
    mfcc_input0 = mfcc[n_s:n_e].reshape((-1, n_times, mfcc.shape[-1]))
    mfcc_input1 = None
    if n_e - n_s > n_times:
        mfcc_input1 = mfcc[(n_s + n_times // 2):(n_e - n_times // 2)].reshape((-1, n_times, mfcc.shape[-1]))
    return mfcc_input0, mfcc_input1, mel[n_s:n_e], stft[n_s:n_e], n_s, n_e


def _zero_pad(*specs, pad_len=0):
    """Sound_DS._zero_pad (sound_ds.py:246-259): rows of float64 zeros appended (np.zeros, so the result is float64)."""
    return [np.concatenate([spec, np.zeros((pad_len, spec.shape[1]))], axis=0) for spec in specs]


def spec_window_sampler(cache, sample_ids, n_timesteps, batch_size=32, n_epochs=1, randomize_samples=True,
                        sample_trn=True, prop_val=0.3, random_seed=None, yield_idxs=False):
    """Literal transcription of Sound_DS.spec_window_sampler (sound_ds.py:262-350) over an open cache
    (``cache[group][str(i)]``); ``sample_ids`` stands for ``np.arange(f_s.shape[0])[f_s]`` of the reader's filter and
    ``random_seed`` for ``self.random_seed``.  Draws from NumPy's global generator exactly like the reference."""
    samples_v = np.array([str(i) for i in sample_ids])
    if prop_val > 0.0:
        np.random.seed(0)                                                  # :270
        idx_v = np.arange(samples_v.shape[0])
        np.random.shuffle(idx_v)
        n_val = int(prop_val * samples_v.shape[0])
        idx_trn = idx_v[:-n_val]
        idx_val = idx_v[-n_val:]
        samples_v = samples_v[idx_trn] if sample_trn else samples_v[idx_val]
        np.random.seed(random_seed)                                        # :284
    mfcc_v, mel_dB_v, power_dB_v, idxs_v = [], [], [], []
    for i_epoch in range(n_epochs):
        if randomize_samples:
            np.random.shuffle(samples_v)                                   # :296
        for i_sample in samples_v:
            spec_len = cache['mfcc'][i_sample].shape[0]
            if spec_len <= n_timesteps:                                    # :301-311
                i_s, i_e = 0, n_timesteps
                mfcc, mel_dB, power_dB = _zero_pad(cache['mfcc'][i_sample][:], cache['mel_dB'][i_sample][:],
                                                   cache['power_dB'][i_sample][:], pad_len=n_timesteps - spec_len)
            else:                                                          # :317-324
                i_s = np.random.randint(0, spec_len - n_timesteps)
                i_e = i_s + n_timesteps
                mfcc = cache['mfcc'][i_sample][i_s:i_e]
                mel_dB = cache['mel_dB'][i_sample][i_s:i_e]
                power_dB = cache['power_dB'][i_sample][i_s:i_e]
            mfcc_v.append(mfcc); mel_dB_v.append(mel_dB); power_dB_v.append(power_dB)
            idxs_v.append([i_s, i_e, int(i_sample)])
            if len(mfcc_v) == batch_size:                                  # :334-350
                out = (np.array(mfcc_v), np.array(mel_dB_v), np.array(power_dB_v))
                assert out[0].shape[1] == out[1].shape[1] == out[2].shape[1] == n_timesteps
                yield out + (np.array(idxs_v),) if yield_idxs else out
                mfcc_v, mel_dB_v, power_dB_v, idxs_v = [], [], [], []


def window_sampler(cache, sample_ids, n_timesteps, batch_size=32, n_epochs=1, randomize_samples=True, yield_idxs=False):
    """Literal transcription of TIMIT.window_sampler (TIMIT_reader.py:474-523): (mfcc window, phn window) batches;
    utterances with ``spec_len <= n_timesteps`` are skipped without drawing a random number."""
    samples_v = [str(i) for i in sample_ids]                               # a list here (:478), an array above
    x_v, y_v, idxs_v = [], [], []
    for i_epoch in range(n_epochs):
        if randomize_samples:
            np.random.shuffle(samples_v)                                   # :489
        for i_sample in samples_v:
            spec_len = cache['mfcc'][i_sample].shape[0]
            if spec_len <= n_timesteps:                                    # :496-497
                continue
            i_s = np.random.randint(0, spec_len - n_timesteps)             # :501
            i_e = i_s + n_timesteps
            x_v.append(cache['mfcc'][i_sample][i_s:i_e])
            y_v.append(cache['phn'][i_sample][i_s:i_e])
            idxs_v.append([i_s, i_e, int(i_sample)])
            if len(x_v) == batch_size:                                     # :511-523
                x, y = np.array(x_v), np.array(y_v)
                assert x.shape[1] == y.shape[1] == n_timesteps
                yield (x, y, np.array(idxs_v)) if yield_idxs else (x, y)
                x_v, y_v, idxs_v = [], [], []
