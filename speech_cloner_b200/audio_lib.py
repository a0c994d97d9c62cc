"""Drop-in for the reference's ``audio_lib`` hot path, computed on a B200 by ``libspeechdsp.so``.

``from speech_cloner_b200.audio_lib import calc_MFCC_input, from_power_to_wav, ...`` replaces
``from audio_lib import ...`` (reference callers: TIMIT_reader.py:10, ARCTIC_reader.py:11,
TARGET_spk_reader.py:11, test.py:23).  Signatures, argument order, defaults and the reference's
spellings (``mfcc_normaleze_first_mfcc``, ``calc_mfcc_derivate``, ``realse``) are kept
(/root/reference/audio_lib.py:12, :31, :51, :89-104, :249, :278-287).

NumPy in -> freshly allocated NumPy out, like the reference.  Additionally every function accepts
CUDA ``torch.Tensor`` inputs and then returns CUDA tensors without touching the host.  PyTorch is
used only for device buffers and streams.  There is no CPU fallback: without the built library or
without a GPU the calls raise.

Extra keyword: ``phase0`` on ``griffin_lim_alg`` / ``from_power_to_wav`` injects the
initial phase; ``None`` draws ``np.pi * np.random.rand(*stft_amp.shape)`` on the host from the
global NumPy state exactly as audio_lib.py:255 does.

Batched entry points (``calc_MFCC_input_batch``, ``from_power_to_wav_batch``) run the readers'
``for i_sample in range(n_samples)`` loops (TIMIT_reader.py:169, ARCTIC_reader.py:134) as one
ragged GPU batch.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Optional, Sequence

import numpy as np
from scipy import signal as _signal

from . import _lib

__all__ = [
    "calc_preemphasis", "calc_inv_preemphasis", "calc_PHN_target", "calc_MFCC_input",
    "griffin_lim_alg", "from_power_to_wav", "calc_MFCC_input_batch", "from_power_to_wav_batch",
    "griffin_lim_batch", "DspPlan", "FrontendLayout", "launch_count", "launch_count_reset",
]


def _torch():
    import torch
    return torch


def _require_cuda():
    torch = _torch()
    if not torch.cuda.is_available():
        raise _lib.SpeechDspError("no CUDA device available: speech_cloner_b200 has no CPU fallback")
    return torch


def _stream_ptr(torch) -> int:
    return int(torch.cuda.current_stream().cuda_stream)


# ----------------------------------------------------------------------------------- plans
class DspPlan:
    """Device-resident constant tables for one DSP parameter set (wraps ``sc_plan``)."""

    _cache: dict = {}
    _lock = threading.Lock()

    def __init__(self, sr=16000, n_fft=400, win_length=400, hop_length=80, n_mels=80, n_mfcc=40,
                 window="hann", pre_emphasis=0.97, mfcc_normaleze_first_mfcc=True, mfcc_norm_factor=0.01,
                 calc_mfcc_derivate=False, M_dB_norm_factor=0.01, P_dB_norm_factor=0.01,
                 mean_abs_amp_norm=0.003, clip_output=True, fft_precision="fp64"):
        lib = _lib.load()
        _require_cuda()
        if n_fft is None:
            n_fft = win_length
        win = _window_array(window, int(win_length))
        prm = _lib.ScParams()
        prm.sample_rate = int(sr); prm.n_fft = int(n_fft); prm.win_length = int(win_length)
        prm.hop_length = int(hop_length); prm.n_mels = int(n_mels); prm.n_mfcc = int(n_mfcc)
        prm.mfcc_normalize_first = int(bool(mfcc_normaleze_first_mfcc))
        prm.calc_mfcc_derivative = int(bool(calc_mfcc_derivate))
        prm.clip_output = int(bool(clip_output))
        if fft_precision not in ("fp64", "fp32"):
            raise ValueError("fft_precision must be 'fp64' or 'fp32'")
        prm.fft_precision = 1 if fft_precision == "fp32" else 0
        prm.pre_emphasis = float(pre_emphasis); prm.mfcc_norm_factor = float(mfcc_norm_factor)
        prm.m_db_norm_factor = float(M_dB_norm_factor); prm.p_db_norm_factor = float(P_dB_norm_factor)
        prm.mean_abs_amp_norm = float(mean_abs_amp_norm)
        self._win = np.ascontiguousarray(win, dtype=np.float64)
        prm.window_host = self._win.ctypes.data_as(C.POINTER(C.c_double))
        handle = C.c_void_p()
        _lib.check(lib.sc_plan_create(C.byref(prm), C.byref(handle)), "sc_plan_create")
        self._h = handle
        self._lib = lib
        self.sr, self.n_fft, self.win_length, self.hop_length = int(sr), int(n_fft), int(win_length), int(hop_length)
        self.n_mels, self.n_mfcc = int(n_mels), int(n_mfcc)
        self.n_bins = 1 + self.n_fft // 2
        self.mfcc_width = self.n_mfcc * (2 if calc_mfcc_derivate else 1)
        self.fast_path = bool(lib.sc_plan_is_fast_path(handle))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.sc_plan_destroy(h)
            except Exception:
                pass

    @classmethod
    def get(cls, **kw) -> "DspPlan":
        """Cached plan (the reference rebuilds the mel / DCT matrices on every call, :160-176)."""
        win = kw.get("window", "hann")
        wkey = win if isinstance(win, (str, tuple, float, int)) else ("arr", np.asarray(win).tobytes())
        torch = _require_cuda()
        # a plan is single-threaded (speechdsp.h): one cached plan per (device, thread, parameter set)
        key = (torch.cuda.current_device(), threading.get_ident(), wkey) + \
            tuple(sorted((k, v) for k, v in kw.items() if k != "window"))
        with cls._lock:
            plan = cls._cache.get(key)
            if plan is None:
                plan = cls._cache[key] = cls(**kw)
        return plan

    def num_frames(self, n_samples: int) -> int:
        return 1 + int(n_samples) // self.hop_length

    def reserve(self, max_samples: int, max_frames: int, max_utts: int) -> None:
        """Pre-size every buffer of the plan so that later calls within these bounds never allocate (``sc_plan_reserve``)."""
        _lib.check(self._lib.sc_plan_reserve(self._h, int(max_samples), int(max_frames), int(max_utts)), "sc_plan_reserve")

    def poll_status(self) -> int:
        """Device-side error flags since the last poll (waits for the current stream); bit 0 = all-zero utterance."""
        flags = C.c_int32(0)
        _lib.check(self._lib.sc_plan_poll_status(self._h, C.byref(flags), _stream_ptr(_torch())), "sc_plan_poll_status")
        return int(flags.value)

    def profile(self, on: bool) -> None:
        """Record CUDA events between the kernels of the next batch call (see ``sc_profile_enable``)."""
        _lib.check(self._lib.sc_profile_enable(self._h, int(bool(on))), "sc_profile_enable")

    def profile_read(self):
        out = (C.c_double * 4)()
        _lib.check(self._lib.sc_profile_read(self._h, out), "sc_profile_read")
        return [float(v) for v in out]


def _window_array(window, win_length: int) -> np.ndarray:
    """librosa 0.6 ``filters.get_window(window, win_length, fftbins=True)`` (audio_lib.py:145)."""
    if callable(window):
        w = np.asarray(window(win_length), dtype=np.float64)
    elif isinstance(window, (str, tuple)) or np.isscalar(window):
        w = _signal.get_window(window, win_length, fftbins=True)
    else:
        w = np.asarray(window, dtype=np.float64)
    if w.shape != (win_length,):
        raise ValueError("window must have win_length samples")
    return w


# ------------------------------------------------------------------------- input validation
def _valid_audio(y):
    """librosa ``util.valid_audio`` as reached through ``stft``."""
    if not isinstance(y, np.ndarray):
        raise ValueError("data must be of type numpy.ndarray")
    if not np.issubdtype(y.dtype, np.floating):
        raise ValueError("data must be floating-point")
    if y.ndim != 1:
        raise ValueError("Invalid shape for monophonic audio: ndim={:d}, shape={}".format(y.ndim, y.shape))
    if y.shape[0] == 0:
        raise ValueError("audio buffer is empty")
    if not np.isfinite(y).all():
        raise ValueError("Audio buffer is not finite everywhere")


def _valid_gain(y, mean_abs_amp_norm):
    """All-zero audio: the reference's gain (audio_lib.py:126) divides by zero and librosa.stft then raises."""
    if mean_abs_amp_norm != 1.0 and not y.any():
        raise ValueError("Audio buffer is not finite everywhere (all-zero audio has no finite gain)")


def _is_tensor(x) -> bool:
    return type(x).__module__.startswith("torch")


def _to_dev_f32(x, torch):
    """1-D/2-D float array or tensor -> contiguous float32 CUDA tensor."""
    if _is_tensor(x):
        return x.to(device="cuda", dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()


def _align(n: int, a: int) -> int:
    return (n + a - 1) // a * a


def _to_host(torch, *tensors):
    """Device tensors -> NumPy arrays through pinned memory (one asynchronous copy each, one synchronisation)."""
    outs = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in tensors]
    for o, t in zip(outs, tensors):
        o.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return tuple(o.numpy() for o in outs)


# ---------------------------------------------------------------------------- front-end
class FrontendLayout:
    """Packed ragged layout of a batch: 16-byte aligned utterance starts in every buffer."""

    def __init__(self, lengths: Sequence[int], hop_length: int):
        self.lengths = [int(n) for n in lengths]
        self.frames = [1 + n // hop_length for n in self.lengths]
        so, fo = [0], [0]
        for n, t in zip(self.lengths, self.frames):
            so.append(so[-1] + _align(n, 4))
            fo.append(fo[-1] + _align(t, 4))
        self.sample_offsets, self.frame_offsets = so, fo
        self.total_samples, self.total_frames = so[-1], fo[-1]
        self.c_sample_offsets = _lib.i64_array(so)
        self.c_sample_lengths = _lib.i64_array(self.lengths)
        self.c_frame_offsets = _lib.i64_array(fo)
        self.c_frame_counts = _lib.i64_array(self.frames)


def frontend_device(plan: DspPlan, wav_dev, layout: FrontendLayout, out=None):
    """Run ``sc_frontend_batch`` on device buffers; returns (mfcc, mel, pdb) packed CUDA tensors."""
    torch = _require_cuda()
    if out is None:
        out = (torch.empty((layout.total_frames, plan.mfcc_width), dtype=torch.float32, device="cuda"),
               torch.empty((layout.total_frames, plan.n_mels), dtype=torch.float32, device="cuda"),
               torch.empty((layout.total_frames, plan.n_bins), dtype=torch.float32, device="cuda"))
    mfcc, mel, pdb = out
    rc = plan._lib.sc_frontend_batch(plan._h, wav_dev.data_ptr(), layout.c_sample_offsets, layout.c_sample_lengths,
                                     len(layout.lengths), mfcc.data_ptr(), mel.data_ptr(), pdb.data_ptr(),
                                     layout.c_frame_offsets, _stream_ptr(torch))
    _lib.check(rc, "sc_frontend_batch")
    return out


FFT_PRECISION = "fp64"
"""Front-end FFT arithmetic used by the reference-signature functions: ``"fp64"`` reproduces the
reference's float64 scipy FFT to the 1e-4/1e-5 tolerance everywhere; ``"fp32"`` is faster but bins
70-80 dB below the utterance maximum can deviate by up to ~5e-5 (documented in DESIGN.md)."""


def mean_abs_device(plan: DspPlan, wav_dev, layout: FrontendLayout):
    """float32 ``np.abs(y).mean()`` per utterance (bit-identical to NumPy), as a CUDA tensor."""
    torch = _require_cuda()
    out = torch.empty(len(layout.lengths), dtype=torch.float32, device="cuda")
    _lib.check(plan._lib.sc_mean_abs_batch(plan._h, wav_dev.data_ptr(), layout.c_sample_offsets,
                                           layout.c_sample_lengths, len(layout.lengths), out.data_ptr(),
                                           _stream_ptr(torch)), "sc_mean_abs_batch")
    return out


def _plan_from_kwargs(sr, pre_emphasis, hop_length, win_length, n_mels, n_mfcc, n_fft, window,
                      mfcc_normaleze_first_mfcc, mfcc_norm_factor, calc_mfcc_derivate, M_dB_norm_factor,
                      P_dB_norm_factor, mean_abs_amp_norm, clip_output, fft_precision=None) -> DspPlan:
    if n_fft is None:
        n_fft = win_length
    return DspPlan.get(sr=sr, n_fft=int(n_fft), win_length=int(win_length), hop_length=int(hop_length),
                       n_mels=int(n_mels), n_mfcc=int(n_mfcc), window=window, pre_emphasis=float(pre_emphasis),
                       mfcc_normaleze_first_mfcc=bool(mfcc_normaleze_first_mfcc),
                       mfcc_norm_factor=float(mfcc_norm_factor), calc_mfcc_derivate=bool(calc_mfcc_derivate),
                       M_dB_norm_factor=float(M_dB_norm_factor), P_dB_norm_factor=float(P_dB_norm_factor),
                       mean_abs_amp_norm=float(mean_abs_amp_norm), clip_output=bool(clip_output),
                       fft_precision=fft_precision or FFT_PRECISION)


def calc_MFCC_input_batch(wavs, sr=16000, pre_emphasis=0.97, hop_length=40, win_length=400, n_mels=128,
                          n_mfcc=40, n_fft=None, window='hann', mfcc_normaleze_first_mfcc=True,
                          mfcc_norm_factor=0.01, calc_mfcc_derivate=False, M_dB_norm_factor=0.01,
                          P_dB_norm_factor=0.01, mean_abs_amp_norm=0.003, clip_output=True, return_device=False,
                          fft_precision=None):
    """``calc_MFCC_input`` over a list of waveforms as ONE ragged GPU batch.

    Returns a list of ``(MFCC, M_dB, P_dB)`` triples (NumPy views of three packed host arrays, or
    CUDA tensor views when ``return_device``), one per utterance, identical to calling the
    reference function in a loop (TIMIT_reader.py:169-190).
    """
    wavs = list(wavs)
    if not wavs:
        return []
    device_in = all(_is_tensor(w) for w in wavs)
    if not device_in:
        for w in wavs:
            _valid_audio(w)                                   # before any device work, like librosa.stft
            _valid_gain(w, mean_abs_amp_norm)
    torch = _require_cuda()
    plan = _plan_from_kwargs(sr, pre_emphasis, hop_length, win_length, n_mels, n_mfcc, n_fft, window,
                             mfcc_normaleze_first_mfcc, mfcc_norm_factor, calc_mfcc_derivate, M_dB_norm_factor,
                             P_dB_norm_factor, mean_abs_amp_norm, clip_output, fft_precision)
    layout = FrontendLayout([int(w.shape[0]) for w in wavs], plan.hop_length)
    if device_in:
        wav_dev = torch.zeros(layout.total_samples, dtype=torch.float32, device="cuda")
        for w, o in zip(wavs, layout.sample_offsets):
            wav_dev[o:o + w.shape[0]] = w
    else:
        host = torch.empty(layout.total_samples, dtype=torch.float32, pin_memory=True)   # torch's caching pinned allocator
        hv = host.numpy()
        for w, o, o1 in zip(wavs, layout.sample_offsets, layout.sample_offsets[1:]):
            hv[o:o + w.shape[0]] = w
            hv[o + w.shape[0]:o1] = 0.0
        wav_dev = host.to("cuda", non_blocking=True)
    mfcc, mel, pdb = frontend_device(plan, wav_dev, layout)
    if device_in and plan.poll_status() & 1:
        raise ValueError("Audio buffer is not finite everywhere (all-zero audio has no finite gain)")
    if not (return_device or device_in):
        mfcc, mel, pdb = _to_host(torch, mfcc, mel, pdb)
    out = []
    for o, t in zip(layout.frame_offsets, layout.frames):
        out.append((mfcc[o:o + t], mel[o:o + t], pdb[o:o + t]))
    return out


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
    """Waveform -> (MFCC (T, n_mfcc[*2]), M_dB (T, n_mels), P_dB (T, 1+n_fft//2)), float32, time-major.

    Same contract as /root/reference/audio_lib.py:89-244.
    """
    res = calc_MFCC_input_batch([y], sr=sr, pre_emphasis=pre_emphasis, hop_length=hop_length,
                                win_length=win_length, n_mels=n_mels, n_mfcc=n_mfcc, n_fft=n_fft, window=window,
                                mfcc_normaleze_first_mfcc=mfcc_normaleze_first_mfcc,
                                mfcc_norm_factor=mfcc_norm_factor, calc_mfcc_derivate=calc_mfcc_derivate,
                                M_dB_norm_factor=M_dB_norm_factor, P_dB_norm_factor=P_dB_norm_factor,
                                mean_abs_amp_norm=mean_abs_amp_norm, clip_output=clip_output)[0]
    if _is_tensor(res[0]):
        return tuple(r.clone() for r in res)
    return tuple(np.ascontiguousarray(r) for r in res)


# -------------------------------------------------------------------- pre- / de-emphasis
def _emphasis(wav, coeff, inverse: bool):
    torch = _require_cuda()
    lib = _lib.load()
    device_in = _is_tensor(wav)
    if not device_in:
        wav = np.asarray(wav)
        if wav.ndim != 1:
            raise ValueError("wav must be one-dimensional")
    # scipy.signal.lfilter computes in float64 whatever the input: float64 input is NOT rounded to float32 first
    f64 = (wav.dtype == torch.float64) if device_in else (wav.dtype == np.float64)
    if f64:
        x = wav.to(device="cuda").contiguous() if device_in else torch.from_numpy(np.ascontiguousarray(wav)).cuda()
        fn, name = (lib.sc_inv_preemphasis_f64, "sc_inv_preemphasis_f64") if inverse else (lib.sc_preemphasis_f64, "sc_preemphasis_f64")
    else:
        x = _to_dev_f32(wav, torch)
        fn, name = (lib.sc_inv_preemphasis, "sc_inv_preemphasis") if inverse else (lib.sc_preemphasis, "sc_preemphasis")
    out = torch.empty(x.shape[0], dtype=torch.float64, device="cuda")
    _lib.check(fn(x.data_ptr(), x.shape[0], float(coeff), out.data_ptr(), _stream_ptr(torch)), name)
    return out if device_in else out.cpu().numpy()


def calc_preemphasis(wav, coeff=0.97):
    """y[n] = x[n] - coeff * x[n-1], zero initial state, float64 (audio_lib.py:12-28)."""
    return _emphasis(wav, coeff, inverse=False)


def calc_inv_preemphasis(preem_wav, coeff=0.97):
    """y[n] = x[n] + coeff * y[n-1], zero initial state, float64 (audio_lib.py:31-47)."""
    return _emphasis(preem_wav, coeff, inverse=True)


def _phn_pick_host(n_samples: int, starts, ends, hop_length: int, win_length: int) -> np.ndarray:
    """Index of the interval calc_PHN_target selects for every frame (audio_lib.py:56-79), on the host."""
    n_frames = int(n_samples / hop_length) + 1
    last = len(starts) - 1
    lo = np.arange(n_frames, dtype=np.int64) * hop_length - win_length // 2
    hi = lo + win_length
    if np.all(np.diff(ends) >= 0):
        # sorted ends: the reference's forward-only cursor is "first interval whose end exceeds lo"
        cur = np.minimum(np.searchsorted(ends, lo, side="right"), last)
    else:
        cur = np.empty(n_frames, dtype=np.int64)
        c = 0
        for t in range(n_frames):
            while ends[c] <= lo[t] and c < last:
                c += 1
            cur[t] = c
    nxt = np.minimum(cur + 1, last)
    ov_cur = np.minimum(ends[cur], hi) - np.maximum(starts[cur], lo)
    ov_nxt = np.minimum(ends[nxt], hi) - np.maximum(starts[nxt], lo)
    return np.where((cur < last) & (ov_cur < ov_nxt), nxt, cur)


def calc_PHN_target(y, phn_v, phn_conv_d, hop_length=40, win_length=400):
    """Per-frame phoneme label by larger window overlap (audio_lib.py:51-85).

    One utterance: integer logic on the host, exactly the reference's cursor.  The dataset-cache builders
    (TIMIT_reader.py:192 and friends) use :func:`calc_PHN_target_batch`, which runs it on the device.
    """
    starts = np.asarray([p[0] for p in phn_v], dtype=np.int64)
    ends = np.asarray([p[1] for p in phn_v], dtype=np.int64)
    pick = _phn_pick_host(int(y.shape[0]), starts, ends, hop_length, win_length)
    return np.array([phn_conv_d[phn_v[i][2]] for i in pick], dtype=np.int32)


def calc_PHN_target_batch(lengths, phn_vs, phn_conv_d, hop_length=40, win_length=400, return_index=False):
    """``calc_PHN_target`` for a batch of utterances in one launch (``sc_phn_target_batch``).

    ``lengths[u]`` is ``y.shape[0]`` of utterance u (or the waveform itself), ``phn_vs[u]`` its list of
    ``(start, end, label)``.  Returns a list of int32 label arrays like the reference's (``phn_conv_d`` applied
    on the host: it may map to one-hot vectors), or the selected interval indices when ``return_index``.
    Interval ends must be non-decreasing inside an utterance (true for .PHN files); otherwise ValueError.
    """
    torch = _require_cuda()
    lens = [int(n.shape[0]) if hasattr(n, "shape") else int(n) for n in lengths]
    if len(lens) != len(phn_vs):
        raise ValueError("lengths and phn_vs differ in length")
    starts, ends, off = [], [], [0]
    for v in phn_vs:
        if len(v) == 0:
            raise ValueError("every utterance needs at least one phoneme interval")
        e = [int(p[1]) for p in v]
        if any(b < a for a, b in zip(e, e[1:])):
            raise ValueError("phoneme interval ends must be non-decreasing (use calc_PHN_target for unsorted lists)")
        starts.extend(int(p[0]) for p in v)
        ends.extend(e)
        off.append(off[-1] + len(v))
    frames = [int(n / hop_length) + 1 for n in lens]
    fo = [0]
    for t in frames:
        fo.append(fo[-1] + t)
    plan = DspPlan.get(n_fft=400, win_length=400, hop_length=80)        # any plan: only its descriptor staging is used
    s_dev = torch.tensor(starts, dtype=torch.int32, device="cuda")
    e_dev = torch.tensor(ends, dtype=torch.int32, device="cuda")
    out = torch.empty(fo[-1], dtype=torch.int32, device="cuda")
    rc = plan._lib.sc_phn_target_batch(plan._h, s_dev.data_ptr(), e_dev.data_ptr(), _lib.i64_array(off),
                                       _lib.i64_array(lens), len(lens), int(hop_length), int(win_length),
                                       out.data_ptr(), _lib.i64_array(fo), _stream_ptr(torch))
    _lib.check(rc, "sc_phn_target_batch")
    idx = out.cpu().numpy()
    res = []
    for u, v in enumerate(phn_vs):
        pick = idx[fo[u]:fo[u + 1]]
        if return_index:
            res.append(pick.copy())
        else:
            table = np.array([phn_conv_d[p[2]] for p in v], dtype=np.int32)
            res.append(table[pick])
    return res


# --------------------------------------------------------------------------- Griffin-Lim
def _gl_plan(win_length, hop_length, n_fft) -> DspPlan:
    if n_fft is None:
        n_fft = win_length
    return DspPlan.get(n_fft=int(n_fft), win_length=int(win_length), hop_length=int(hop_length))


def _freq_major_to_dev(x, n_bins: int, torch, lib):
    """(n_bins, T) float32/float64 array or tensor -> [T][n_bins] float32 CUDA tensor."""
    if _is_tensor(x):
        src = x.to(device="cuda").contiguous()
        if src.dtype not in (torch.float32, torch.float64):
            src = src.to(torch.float32)
    else:
        a = np.ascontiguousarray(x)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float32)
        src = torch.from_numpy(a).cuda()
    if src.ndim != 2 or src.shape[0] != n_bins:
        raise ValueError("expected a (1 + n_fft//2, T) array, got {}".format(tuple(src.shape)))
    T = src.shape[1]
    dst = torch.empty((T, n_bins), dtype=torch.float32, device="cuda")
    _lib.check(lib.sc_transpose_to_f32(src.data_ptr(), int(src.dtype == torch.float64), n_bins, T,
                                       dst.data_ptr(), _stream_ptr(torch)), "sc_transpose_to_f32")
    return dst


class _GlLayout:
    def __init__(self, frames: Sequence[int], hop: int):
        self.frames = [int(t) for t in frames]
        self.samples = [hop * (t - 1) for t in self.frames]
        fo, so = [0], [0]
        for t, n in zip(self.frames, self.samples):
            fo.append(fo[-1] + _align(t, 4))
            so.append(so[-1] + _align(max(n, 1), 4))
        self.frame_offsets, self.sample_offsets = fo, so
        self.c_frame_offsets = _lib.i64_array(fo)
        self.c_frame_counts = _lib.i64_array(self.frames)
        self.c_sample_offsets = _lib.i64_array(so)
        self.c_sample_lengths = _lib.i64_array(self.samples)


def griffin_lim_device(plan: DspPlan, amp_dev, phase0_dev, layout: "_GlLayout", n_iters: int, rms=None, wav_out=None):
    """``sc_griffinlim_batch`` on packed time-major device buffers; returns the packed float32 waveforms."""
    torch = _require_cuda()
    if wav_out is None:
        wav_out = torch.empty(layout.sample_offsets[-1], dtype=torch.float32, device="cuda")
    rc = plan._lib.sc_griffinlim_batch(plan._h, amp_dev.data_ptr(), phase0_dev.data_ptr(), layout.c_frame_offsets,
                                       layout.c_frame_counts, len(layout.frames), int(n_iters), wav_out.data_ptr(),
                                       layout.c_sample_offsets, rms.data_ptr() if rms is not None else None,
                                       _stream_ptr(torch))
    _lib.check(rc, "sc_griffinlim_batch")
    return wav_out


def _print_rms(rms_row):
    for i in range(1, len(rms_row)):
        print(' i={}  mrse_delta = {}'.format(i, rms_row[i]))


def griffin_lim_alg(stft_amp, win_length, hop_length, num_iters=300, n_fft=None, verbose=True, phase0=None):
    """Griffin-Lim from a (1+n_fft//2, T) magnitude; returns float32 (hop*(T-1),) (audio_lib.py:249-274)."""
    if int(num_iters) < 1:
        return None                                                   # the reference's loop never runs: `wav = None` (:252)
    torch = _require_cuda()
    lib = _lib.load()
    plan = _gl_plan(win_length, hop_length, n_fft)
    device_in = _is_tensor(stft_amp)
    shape = tuple(stft_amp.shape)
    if len(shape) != 2 or shape[0] != plan.n_bins:
        raise ValueError("stft_amp must have shape (1 + n_fft//2, T)")
    if shape[1] < 2:
        raise ValueError("stft_amp needs at least 2 frames")
    if phase0 is None:
        phase0 = np.pi * np.random.rand(*shape)                       # :255, global NumPy state
    amp = _freq_major_to_dev(stft_amp, plan.n_bins, torch, lib)
    ph = _freq_major_to_dev(phase0, plan.n_bins, torch, lib)
    layout = _GlLayout([shape[1]], plan.hop_length)
    rms = torch.zeros((1, int(num_iters)), dtype=torch.float32, device="cuda") if verbose else None
    wav = griffin_lim_device(plan, amp, ph, layout, int(num_iters), rms)[: layout.samples[0]]
    if verbose:
        _print_rms(rms[0].cpu().numpy())
    return wav.clone() if device_in else wav.cpu().numpy()


def griffin_lim_batch(amps, win_length, hop_length, num_iters=300, n_fft=None, phase0s=None, return_device=False):
    """``griffin_lim_alg`` over a list of TIME-MAJOR (T, bins) magnitudes as one ragged GPU batch."""
    torch = _require_cuda()
    plan = _gl_plan(win_length, hop_length, n_fft)
    amps = list(amps)
    layout = _GlLayout([a.shape[0] for a in amps], plan.hop_length)
    amp_dev = torch.zeros((layout.frame_offsets[-1], plan.n_bins), dtype=torch.float32, device="cuda")
    ph_dev = torch.zeros_like(amp_dev)
    for i, (a, o) in enumerate(zip(amps, layout.frame_offsets)):
        amp_dev[o:o + a.shape[0]] = _to_dev_f32(a, torch)
        p = phase0s[i] if phase0s is not None else np.pi * np.random.rand(plan.n_bins, a.shape[0]).T
        ph_dev[o:o + a.shape[0]] = _to_dev_f32(p, torch)
    wav = griffin_lim_device(plan, amp_dev, ph_dev, layout, int(num_iters))
    outs = [wav[o:o + n] for o, n in zip(layout.sample_offsets, layout.samples)]
    return outs if return_device else [w.cpu().numpy() for w in outs]


_PACK_POOL = None


def _pack_pool():
    """Small thread pool for packing host arrays into pinned staging buffers (NumPy copies release the GIL)."""
    global _PACK_POOL
    if _PACK_POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        _PACK_POOL = ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 2) // 2)))
    return _PACK_POOL


def _stage_rows(items, layout, n_bins, torch, transpose_from_freq_major: bool, lib):
    """List of per-utterance arrays / tensors -> one packed time-major float32 CUDA buffer [rows][n_bins].

    Host arrays travel through ONE pinned staging buffer: groups of utterances are packed by a few threads and each
    group's asynchronous H2D copy (and device transposes) is queued as soon as it is packed, so packing, PCIe and the
    transposes overlap (the per-utterance ``.cuda()`` calls of round 1 cost more than the Griffin-Lim prologue they
    fed).  ``transpose_from_freq_major``: the items are (n_bins, T) in the reference's orientation (float32 or
    float64) and are transposed on the device.
    """
    rows = layout.frame_offsets[-1]
    dst = torch.zeros((rows, n_bins), dtype=torch.float32, device="cuda")
    st = _stream_ptr(torch)
    if all(_is_tensor(x) for x in items):
        for x, o, t in zip(items, layout.frame_offsets, layout.frames):
            if transpose_from_freq_major:
                dst[o:o + t] = x.to(device="cuda", dtype=torch.float32).t()
            else:
                dst[o:o + t] = x.to(device="cuda", dtype=torch.float32)
        return dst
    arrs = [x.cpu().numpy() if _is_tensor(x) else np.asarray(x) for x in items]
    n = len(arrs)
    groups = [(g, min(n, g + max(1, -(-n // 8)))) for g in range(0, n, max(1, -(-n // 8)))]
    pool = _pack_pool()
    if not transpose_from_freq_major:
        host = torch.empty((rows, n_bins), dtype=torch.float32, pin_memory=True)
        hv = host.numpy()

        def pack(i):
            o, o1, t = layout.frame_offsets[i], layout.frame_offsets[i + 1], layout.frames[i]
            hv[o:o + t] = arrs[i]
            hv[o + t:o1] = 0.0
        for g0, g1 in groups:
            list(pool.map(pack, range(g0, g1)))
            r0, r1 = layout.frame_offsets[g0], layout.frame_offsets[g1]
            dst[r0:r1].copy_(host[r0:r1], non_blocking=True)
        return dst
    f64 = any(a.dtype == np.float64 for a in arrs)
    dt_np, dt_t = (np.float64, torch.float64) if f64 else (np.float32, torch.float32)
    offs = [0]
    for t in layout.frames:
        offs.append(offs[-1] + t * n_bins)
    host = torch.empty(offs[-1], dtype=dt_t, pin_memory=True)
    src = torch.empty(offs[-1], dtype=dt_t, device="cuda")
    hv = host.numpy()

    def pack_t(i):
        hv[offs[i]:offs[i + 1]].reshape(n_bins, layout.frames[i])[...] = arrs[i]
    for g0, g1 in groups:
        list(pool.map(pack_t, range(g0, g1)))
        src[offs[g0]:offs[g1]].copy_(host[offs[g0]:offs[g1]], non_blocking=True)
        for i in range(g0, g1):
            _lib.check(lib.sc_transpose_to_f32(src.data_ptr() + offs[i] * src.element_size(), int(f64), n_bins, layout.frames[i],
                                               dst.data_ptr() + layout.frame_offsets[i] * n_bins * 4, st), "sc_transpose_to_f32")
    return dst


def _power_to_wav_device(plan, lib, torch, Ps, phase0s, P_dB_norm_factor, pre_emphasis, mean_abs_amp_norm, n_iter, realse, verbose):
    """Stage, convert, invert and de-emphasise one group of spectrograms; everything is queued on the current stream.
    Returns (packed float64 device waveforms, layout)."""
    layout = _GlLayout([P.shape[0] for P in Ps], plan.hop_length)
    n = len(Ps)
    p_dev = _stage_rows(Ps, layout, plan.n_bins, torch, False, lib)
    ph_dev = _stage_rows(list(phase0s), layout, plan.n_bins, torch, True, lib)
    st = _stream_ptr(torch)
    _lib.check(lib.sc_power_to_amp_batch(plan._h, p_dev.data_ptr(), layout.c_frame_offsets, layout.c_frame_counts,
                                         n, float(P_dB_norm_factor), float(realse), p_dev.data_ptr(), st),
               "sc_power_to_amp_batch")                                              # in place (the header allows aliasing)
    rms = torch.zeros((n, int(n_iter)), dtype=torch.float32, device="cuda") if verbose else None
    wav = griffin_lim_device(plan, p_dev, ph_dev, layout, int(n_iter), rms)
    if verbose:
        _print_rms(rms[0].cpu().numpy())
    out = torch.empty(layout.sample_offsets[-1], dtype=torch.float64, device="cuda")
    _lib.check(lib.sc_deemph_renorm_batch(plan._h, wav.data_ptr(), layout.c_sample_offsets, layout.c_sample_lengths,
                                          n, float(pre_emphasis), float(mean_abs_amp_norm), out.data_ptr(), st),
               "sc_deemph_renorm_batch")
    return out, layout


def from_power_to_wav_batch(Ps, P_dB_norm_factor=0.01, pre_emphasis=0.97, hop_length=40, win_length=800,
                            mean_abs_amp_norm=0.01, n_iter=200, n_fft=None, realse=1.0, verbose=False,
                            phase0s=None, return_device=False, pipeline_groups=4):
    """``from_power_to_wav`` over a list of (T, bins) spectrograms as one ragged GPU batch.

    ``phase0s`` is a list of (bins, T) initial phases (reference orientation) or ``None`` (drawn per utterance from the
    global NumPy state in list order, like calling the reference in a loop).  Host inputs are processed in
    ``pipeline_groups`` groups queued back to back on the current stream: while the GPU inverts group g the host packs
    group g+1 into pinned memory, so the staging cost is hidden behind the 200 iterations (utterances are independent:
    the results do not depend on the grouping).
    """
    if int(n_iter) < 1:
        raise ValueError("n_iter must be >= 1 (the reference's griffin_lim_alg returns None for 0 iterations)")
    torch = _require_cuda()
    lib = _lib.load()
    plan = _gl_plan(win_length, hop_length, n_fft)
    Ps = list(Ps)
    for P in Ps:
        if len(P.shape) != 2 or P.shape[1] != plan.n_bins:
            raise ValueError("P must have shape (T, 1 + n_fft//2)")
        if P.shape[0] < 2:
            raise ValueError("P needs at least 2 frames")
    n = len(Ps)
    if phase0s is None:
        phase0s = [np.pi * np.random.rand(plan.n_bins, P.shape[0]) for P in Ps]      # :255
    phase0s = list(phase0s)
    on_host = not all(_is_tensor(x) for x in Ps + phase0s)
    # groups of at least 8 spectrograms (a group should still fill the GPU: ~300 tiles of 28 hops)
    g = max(1, min(int(pipeline_groups), n // 8)) if (on_host and not verbose) else 1
    step = -(-n // g)
    bounds = [(i, min(n, i + step)) for i in range(0, n, step)]
    kw = dict(P_dB_norm_factor=P_dB_norm_factor, pre_emphasis=pre_emphasis, mean_abs_amp_norm=mean_abs_amp_norm,
              n_iter=n_iter, realse=realse, verbose=verbose)
    parts = [_power_to_wav_device(plan, lib, torch, Ps[a:b], phase0s[a:b], **kw) for a, b in bounds]
    outs = []
    if return_device:
        for out, layout in parts:
            outs += [out[o:o + m] for o, m in zip(layout.sample_offsets, layout.samples)]
        return outs
    # librosa.stft inside the reference's loop (audio_lib.py:267) rejects a non-finite waveform: NaN / Inf input maps, or an
    # all-zero map with realse != 1 (0 / 0 at :296), raise there from the second iteration on
    finite = torch.stack([torch.isfinite(out).all() for out, _ in parts]).all() if int(n_iter) >= 2 else None
    hosts = _to_host(torch, *[out for out, _ in parts])
    if finite is not None and not bool(finite):
        raise ValueError("Audio buffer is not finite everywhere")
    for host, (_, layout) in zip(hosts, parts):
        outs += [host[o:o + m] for o, m in zip(layout.sample_offsets, layout.samples)]
    return outs


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
    """Normalised power-dB map (T, 1+n_fft//2) -> waveform float64 (hop*(T-1),) (audio_lib.py:278-308)."""
    device_in = _is_tensor(P)
    res = from_power_to_wav_batch([P], P_dB_norm_factor=P_dB_norm_factor, pre_emphasis=pre_emphasis,
                                  hop_length=hop_length, win_length=win_length, mean_abs_amp_norm=mean_abs_amp_norm,
                                  n_iter=n_iter, n_fft=n_fft, realse=realse, verbose=verbose,
                                  phase0s=None if phase0 is None else [phase0], return_device=device_in)[0]
    if device_in:
        return res.clone()
    # without de-emphasis the reference never leaves float32 (:304-306)
    return res.astype(np.float32) if pre_emphasis == 0 else res


def launch_count() -> int:
    """Kernels launched by the library since the last reset (bench.py's ``gpu_launches``)."""
    return int(_lib.load().sc_launch_count())


def launch_count_reset() -> None:
    _lib.load().sc_launch_count_reset()


# ------------------------------------------------------------------ host-buffer pipeline
class FrontendPipeline:
    """Featurise a fixed ragged layout from HOST buffers: pinned H2D -> kernels -> pinned D2H.

    This is the dataset-builder path (the readers' per-utterance loop, TIMIT_reader.py:169-201, as
    one call): the batch is cut into ``n_chunks`` utterance groups that rotate over ``n_streams``
    CUDA streams, so the upload of chunk c+1 and the download of chunk c-1 overlap the kernels of
    chunk c.  Each stream owns its own plan (workspace).  Results land in three packed pinned host
    arrays; ``views()`` gives the per-utterance (MFCC, M_dB, P_dB) triples.
    """

    def __init__(self, lengths: Sequence[int], n_chunks: int = 8, n_streams: int = 3, ramp: bool = True, **plan_kw):
        torch = _require_cuda()
        self.torch = torch
        self.plans = [DspPlan(**plan_kw) for _ in range(n_streams)]
        p0 = self.plans[0]
        self.layout = FrontendLayout(lengths, p0.hop_length)
        lay = self.layout
        n = len(lay.lengths)
        n_chunks = max(1, min(n_chunks, n))
        # contiguous utterance groups: the first ones are small and double in size (the download engine, which
        # bounds the whole pass, starts after ~1/64 of the batch instead of after 1/n_chunks), the rest are equal
        fracs, f = [], 1.0 / (8 * n_chunks) if ramp else 1.0 / n_chunks
        while ramp and f < 1.0 / n_chunks and sum(fracs) + f < 1.0:
            fracs.append(f)
            f *= 2
        rest = 1.0 - sum(fracs)
        k = max(1, int(round(rest * n_chunks)))
        fracs += [rest / k] * k
        cuts, c = [], 0.0
        for fr in fracs[:-1]:
            c += fr
            cuts.append(c * lay.total_frames)
        bounds, acc = [0], 0
        for i, t in enumerate(lay.frames):
            acc += _align(t, 4)
            if len(bounds) - 1 < len(cuts) and acc >= cuts[len(bounds) - 1] and i + 1 < n:
                bounds.append(i + 1)
        bounds.append(n)
        self.chunks = []
        for a, b in zip(bounds[:-1], bounds[1:]):
            sub = FrontendLayout(lay.lengths[a:b], p0.hop_length)
            self.chunks.append((a, b, sub, lay.sample_offsets[a], lay.frame_offsets[a]))
        self.streams = [torch.cuda.Stream() for _ in range(n_streams)]
        self.wav_host = torch.zeros(lay.total_samples, dtype=torch.float32).pin_memory()
        self.wav_dev = torch.empty(lay.total_samples, dtype=torch.float32, device="cuda")
        shapes = ((lay.total_frames, p0.mfcc_width), (lay.total_frames, p0.n_mels), (lay.total_frames, p0.n_bins))
        self.out_dev = tuple(torch.empty(s, dtype=torch.float32, device="cuda") for s in shapes)
        self.out_host = tuple(torch.empty(s, dtype=torch.float32).pin_memory() for s in shapes)
        self.h2d_bytes = sum(c[2].total_samples for c in self.chunks) * 4
        self.d2h_bytes = sum(c[2].total_frames for c in self.chunks) * 4 * (p0.mfcc_width + p0.n_mels + p0.n_bins)

    def load(self, wavs) -> None:
        """Pack host waveforms into the pinned staging buffer (not part of ``run``)."""
        hv = self.wav_host.numpy()
        for w, o in zip(wavs, self.layout.sample_offsets):
            hv[o:o + len(w)] = w

    def run(self):
        """One pass: H2D, kernels, D2H for every chunk; returns after the last byte is on the host."""
        torch = self.torch
        cur = torch.cuda.current_stream()
        for s in self.streams:
            s.wait_stream(cur)
        for i, (a, b, sub, s_off, f_off) in enumerate(self.chunks):
            st = self.streams[i % len(self.streams)]
            plan = self.plans[i % len(self.plans)]
            with torch.cuda.stream(st):
                ns, nf = sub.total_samples, sub.total_frames
                self.wav_dev[s_off:s_off + ns].copy_(self.wav_host[s_off:s_off + ns], non_blocking=True)
                outs = tuple(o[f_off:f_off + nf] for o in self.out_dev)
                frontend_device(plan, self.wav_dev[s_off:s_off + ns], sub, outs)
                for h, d in zip(self.out_host, outs):
                    h[f_off:f_off + nf].copy_(d, non_blocking=True)
        for s in self.streams:
            cur.wait_stream(s)
        cur.synchronize()
        return self.out_host

    def views(self):
        lay = self.layout
        hs = [h.numpy() for h in self.out_host]
        return [tuple(h[o:o + t] for h in hs) for o, t in zip(lay.frame_offsets, lay.frames)]
