"""Randomised parity soak: ragged front-end batches and short Griffin-Lim inversions against the CPU oracle for a fixed time.

    python scripts/soak.py [seconds] [seed]

Lengths are drawn around the places where the kernels switch paths: 1 .. n_fft/2 (multiple reflections), the 4 496-sample
threshold of the warp-specialised pass A, 24-frame and 28-frame tile edges, the 8 000-sample sub-trees of the |y| sum,
multiples of hop +- 1.  Tolerances are the north-star ones (1e-4 relative + 1e-5 absolute; SNR >= 40 dB).  Test
infrastructure: imports the oracle as the checker.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import audio_lib_oracle as oracle            # noqa: E402
from speech_cloner_b200 import audio_lib as al, synth    # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
HP = dict(synth.HP_ENC)
ANCHORS = [201, 400, 401, 2240, 4495, 4496, 4497, 8000, 8001, 16000, 24 * 80, 28 * 80, 48 * 80 + 200, 64000]


def rand_len():
    kind = rng.integers(0, 5)
    if kind == 0:
        return int(rng.integers(2, 400))
    if kind == 1:
        return int(rng.integers(400, 6000))
    if kind == 2:
        return int(max(2, ANCHORS[rng.integers(0, len(ANCHORS))] + rng.integers(-3, 4)))
    if kind == 3:
        return int(80 * rng.integers(5, 400) + rng.integers(-1, 2))
    return int(rng.integers(6000, 70000))


def rand_wave(n):
    y = synth.utterance(int(rng.integers(0, 1 << 30)), max(n, 16) / 16000.0)[:n].copy()
    if len(y) < n:
        y = np.concatenate([y, 0.01 * rng.standard_normal(n - len(y)).astype(np.float32)])
    mode = rng.integers(0, 6)
    if mode == 0:
        y *= np.float32(10.0 ** rng.uniform(-3, 2))
    elif mode == 1:
        y[: len(y) // 2] = 0.0                                  # half silence
    elif mode == 2:
        y = (0.1 * np.sin(2 * np.pi * rng.uniform(50, 7900) * np.arange(n) / 16000.0)).astype(np.float32)   # pure tone
    elif mode == 3:
        y = (0.05 * rng.standard_normal(n)).astype(np.float32)  # white noise: small dynamic range (true min above the floor)
    if not np.any(y):
        y[0] = 1e-3
    return np.ascontiguousarray(y, dtype=np.float32)


def close(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return a.shape == b.shape and bool(np.all(np.abs(a - b) <= 1e-5 + 1e-4 * np.abs(b)))


t_end = time.time() + budget
n_batches = n_utts = n_gl = 0
worst = 0.0
while time.time() < t_end:
    kw = dict(HP)
    if rng.random() < 0.2:
        kw.update(calc_mfcc_derivate=bool(rng.integers(0, 2)), clip_output=bool(rng.integers(0, 2)),
                  mfcc_normaleze_first_mfcc=bool(rng.integers(0, 2)))
    lens = [rand_len() for _ in range(int(rng.integers(1, 9)))]
    if kw["calc_mfcc_derivate"]:
        lens = [max(n, 80) for n in lens]                         # the delta needs two frames (reference: shape error)
    wavs = [rand_wave(n) for n in lens]
    got = al.calc_MFCC_input_batch(wavs, **kw)
    for y, g in zip(wavs, got):
        try:
            want = oracle.calc_MFCC_input(y, **kw)
        except ValueError as exc:
            print(f"ORACLE RAISES ({exc}) where the product returned: len {len(y)} max|y| {np.abs(y).max():.3e} mean|y| "
                  f"{np.abs(y).mean():.3e} finite product output {all(bool(np.isfinite(a).all()) for a in g)}")
            np.save("gpurun_out/soak_fail_wav.npy", y)
            sys.exit(1)
        for name, a, b in zip(("mfcc", "mel", "pdb"), g, want):
            if not close(a, b):
                err = np.abs(a.astype(np.float64) - b)
                print(f"MISMATCH {name}: len {len(y)} lens {lens} kw-diff {[k for k in kw if kw[k] != HP[k]]} max err {err.max():.3e}")
                np.save("gpurun_out/soak_fail_wav.npy", y)
                sys.exit(1)
            worst = max(worst, float(np.max(np.abs(a.astype(np.float64) - b))))
    n_batches += 1; n_utts += len(wavs)
    if n_batches % 4 == 0:                                        # a short Griffin-Lim from one of the maps
        P = got[int(np.argmax(lens))][2]
        T = int(min(P.shape[0], rng.integers(2, 120)))
        if T >= 2:
            n_it = int(rng.integers(1, 8))
            ph = np.pi * rng.random((201, T))
            kwg = dict(P_dB_norm_factor=0.01, pre_emphasis=0.97, hop_length=80, win_length=400, mean_abs_amp_norm=0.045,
                       n_iter=n_it, verbose=False, phase0=ph, realse=float(rng.choice([1.0, 1.2])))
            a = b = err_a = err_b = None
            try:
                a = al.from_power_to_wav(P[:T], **kwg)
            except ValueError as exc:
                err_a = exc
            try:
                with np.errstate(all="ignore"):
                    b = oracle.from_power_to_wav(P[:T], **kwg)
            except ValueError as exc:
                err_b = exc
            if (err_a is None) != (err_b is None):
                print(f"ERROR PARITY in Griffin-Lim: product {err_a!r} oracle {err_b!r}; T {T} iterations {n_it} realse "
                      f"{kwg['realse']} P range {P[:T].min():.3e}..{P[:T].max():.3e}")
                np.save("gpurun_out/soak_fail_P.npy", P[:T]); np.save("gpurun_out/soak_fail_ph.npy", ph)
                sys.exit(1)
            if err_a is not None:
                n_gl += 1
                continue
            if not np.isfinite(b).all():                          # one iteration on a degenerate map: NaN in, NaN out, both sides
                if not (a.shape == b.shape and np.array_equal(np.isfinite(a), np.isfinite(b))):
                    print(f"MISMATCH Griffin-Lim (non-finite pattern): T {T} iterations {n_it}")
                    sys.exit(1)
                n_gl += 1
                continue
            snr = 10 * np.log10(np.sum(b ** 2) / max(np.sum((a - b) ** 2), 1e-300))
            if not (a.shape == b.shape and snr >= 40.0):
                print(f"MISMATCH Griffin-Lim: T {T} iterations {n_it} SNR {snr:.1f} dB")
                sys.exit(1)
            n_gl += 1
print(f"soak ok: {n_batches} ragged batches, {n_utts} utterances, {n_gl} Griffin-Lim inversions in {budget:.0f} s; "
      f"worst absolute feature error {worst:.2e}; seed {seed}")
