"""Conversion glue either side of Griffin-Lim: the window / half-offset batching and zero
padding in front of ``decoder.predict`` (test.py:92-128), ``compound`` stitching of the decoder's window predictions
(test.py:46-84) and the peak normalisation of ``write_wav(norm=True)`` (test.py:177-179).

Both keep the data where it is: CUDA tensors in -> CUDA tensors out, so a decoder's device output can be stitched,
inverted by ``from_power_to_wav`` and normalised without a host round trip; NumPy in -> NumPy out reproduces the
reference's slicing.  No arithmetic happens here except one division by the peak.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from . import audio_lib as al


def window_batches(mfcc, mel, stft, cfg_d, t_s=5, t_e=60):
    """The batching ``conversion2`` does before it calls the decoder (test.py:92-128), on NumPy arrays or torch tensors
    (CUDA tensors stay on the device: everything is views plus one zero pad).

    Pads the three time-major feature maps with zero rows to a multiple of ``n_timesteps``, cuts frames
    ``[n_s, n_e)`` (``t_s`` / ``t_e`` seconds, a whole number of windows), and returns a dict with
    ``mfcc_input0`` ``(N, n_timesteps, F)`` (the aligned windows), ``mfcc_input1`` ``(N-1, n_timesteps, F)`` (the same
    frames shifted by half a window, ``None`` when there is only one window), ``mel_true`` / ``stft_true`` (the frames
    the prediction is compared with) and ``n_s`` / ``n_e``.  Raises like the reference when ``n_e <= n_s``.
    """
    hop, n_times = int(cfg_d['hop_length']), int(cfg_d['n_timesteps'])
    is_t = al._is_tensor(mfcc)
    if mfcc.shape[0] % n_times != 0:
        pad_len = n_times - (mfcc.shape[0] % n_times)

        def pad(x):
            if is_t:
                import torch
                return torch.cat([x, torch.zeros((pad_len, x.shape[1]), dtype=x.dtype, device=x.device)], dim=0)
            return np.concatenate([x, np.zeros((pad_len, x.shape[1]))], axis=0)   # float64 zeros promote like the reference
        mfcc, mel, stft = pad(mfcc), pad(mel), pad(stft)
    n_hop_s = t_s * cfg_d['sample_rate'] // hop
    n_hop_e = min(t_e * cfg_d['sample_rate'] // hop, mfcc.shape[0])
    n_delta = n_times * ((n_hop_e - n_hop_s) // n_times)
    n_s, n_e = n_hop_s, n_hop_s + n_delta
    if n_e <= n_s:
        raise Exception(' - ERROR, translate: n_e <= n_s.')
    out = {"n_s": n_s, "n_e": n_e,
           "mfcc_input0": mfcc[n_s:n_e].reshape((-1, n_times, mfcc.shape[-1])),
           "mfcc_input1": None,
           "mel_true": mel[n_s:n_e], "stft_true": stft[n_s:n_e]}
    if n_e - n_s > n_times:
        out["mfcc_input1"] = mfcc[(n_s + n_times // 2):(n_e - n_times // 2)].reshape((-1, n_times, mfcc.shape[-1]))
    return out


def stitch_predictions(pred0, pred1=None):
    """``compound`` when the half-offset batch exists, else the plain reshape of test.py:134-138."""
    if pred1 is not None:
        return compound(pred0, pred1)
    return pred0.reshape((-1, pred0.shape[-1]))


def compound_segments(n0: int, n1: int, T: int) -> List[Tuple[int, int, int, int]]:
    """The slices ``compound`` concatenates, as (source 0|1, window index, start, stop): the reference loop
    (test.py:54-79) unrolled.  First window of y0 keeps [0, T - T//4), then the middle halves of y1[0], y0[1],
    y1[1], ... while either has windows left, and the last window of y0 contributes [T//4, T)."""
    q = T // 4
    mid_stop = T - q if q else 0                       # x[q:-q] and x[:-q] are EMPTY for q == 0, like the reference
    segs = [(0, 0, 0, mid_stop)]
    i0, i1 = 1, 0
    while True:
        stop = True
        if i1 < n1:
            segs.append((1, i1, q, mid_stop))
            i1 += 1
            stop = False
        if i0 < n0 - 1:
            segs.append((0, i0, q, mid_stop))
            i0 += 1
            stop = False
        if stop:
            break
    segs.append((0, n0 - 1, q, T))
    return segs


def compound(y0, y1):
    """``compound(y0, y1)`` of test.py:46-84 for NumPy arrays or torch tensors (any device): (N, T, X) and
    (N-1, T, X) -> (rows, X)."""
    if y0.ndim != 3 or y1.ndim != 3 or y0.shape[1:] != y1.shape[1:]:
        raise ValueError("compound: y0 (N, T, X) and y1 (M, T, X) must agree in T and X")
    if y0.shape[0] < 1:
        raise ValueError("compound: y0 needs at least one window")
    src = (y0, y1)
    parts = [src[s][i, a:b, :] for s, i, a, b in compound_segments(y0.shape[0], y1.shape[0], y0.shape[1])]
    if al._is_tensor(y0):
        import torch
        return torch.cat(parts, dim=0)
    return np.concatenate(parts, axis=0)


def normalize_wav(y):
    """``librosa.util.normalize(y, norm=inf)`` as applied by ``write_wav(..., norm=True)`` (test.py:177-179)."""
    if al._is_tensor(y):
        import torch
        mag = y.abs().max() if y.numel() else torch.zeros((), dtype=y.dtype, device=y.device)
        tiny = torch.finfo(y.dtype).tiny
        return torch.where(mag < tiny, y, y / torch.clamp(mag, min=tiny))
    y = np.asarray(y)
    mag = np.max(np.abs(y)) if y.size else 0.0
    tiny = np.finfo(y.dtype if np.issubdtype(y.dtype, np.floating) else np.float32).tiny
    return y if mag < tiny else y / mag


def render_windows(p0, p1, normalize=True, **power_to_wav_kw):
    """Stitch window predictions of the power-dB head and render them: ``compound`` -> ``from_power_to_wav``
    (-> peak normalisation), the sequence of test.py:140-179.  Device tensors stay on the device."""
    P = compound(p0, p1)
    y = al.from_power_to_wav(P, **power_to_wav_kw)
    return normalize_wav(y) if normalize else y
