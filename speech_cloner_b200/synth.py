"""Seeded synthetic speech-shaped audio (SURVEY.md §8(d) "Synthetic inputs").

Harmonic stack (24 partials, 1/k roll-off, f0 = 120 +- 40 Hz with a slow vibrato) times a 3 Hz
syllabic envelope, plus a white-noise floor (sigma ~ 0.002), peak ~ 0.1, float32.  Used by the
parity tests, ``bench.py`` and ``__graft_entry__.smoke()``; there is no network for real datasets.
"""
from __future__ import annotations

import numpy as np

HP_ENC = dict(sr=16000, pre_emphasis=0.97, hop_length=80, win_length=400, n_mels=80, n_mfcc=40, n_fft=None,
              window="hann", mfcc_normaleze_first_mfcc=True, mfcc_norm_factor=0.01, calc_mfcc_derivate=True,
              M_dB_norm_factor=0.01, P_dB_norm_factor=0.01, mean_abs_amp_norm=0.003, clip_output=True)
"""calc_MFCC_input kwargs of hp/ds_enc_cfg_d.json == hp/ds_dec_cfg_d.json (hop/win derived from
hop_length_ms 5.0 / win_length_ms 25.0 at 16 kHz, TIMIT_reader.py:20-26)."""


def utterance(seed: int, seconds: float, sr: int = 16000, ds_norm=(0.0, 1.0)) -> np.ndarray:
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    t = np.arange(n, dtype=np.float64) / sr
    f0 = 120.0 + 40.0 * (2.0 * rng.random() - 1.0)
    vib = 1.0 + 0.03 * np.sin(2 * np.pi * (4.0 + rng.random()) * t + 2 * np.pi * rng.random())
    phase = 2 * np.pi * np.cumsum(f0 * vib) / sr
    y = np.zeros(n)
    for k in range(1, 25):
        if k * f0 * 1.03 < sr / 2:
            y += np.sin(k * phase + 2 * np.pi * rng.random()) / k
    env = 0.5 * (1.0 - np.cos(2 * np.pi * 3.0 * t + 2 * np.pi * rng.random()))
    y *= env ** 2
    y *= 0.1 / max(np.abs(y).max(), 1e-12)
    y += 0.002 * rng.standard_normal(n)
    add, mult = ds_norm                       # sound_ds.py:56-63: wav <- mult * (wav + add)
    return (mult * (y + add)).astype(np.float32)


def batch(config_index: int, n_utts: int, seconds: float, sr: int = 16000, ds_norm=(0.0, 1.0)):
    return [utterance(config_index * 1000 + i, seconds, sr, ds_norm) for i in range(n_utts)]
