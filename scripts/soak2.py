"""Second randomised soak (see soak.py): generic STFT geometries, frame labels, emphasis filters and the time-chunked
Griffin-Lim path (emulated ranks, bit-identity with the unchunked call) for a fixed time.

    python scripts/soak2.py [seconds] [seed]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import audio_lib_oracle as oracle            # noqa: E402
from speech_cloner_b200 import audio_lib as al, synth    # noqa: E402
from tests import test_gpu_griffinlim as tg              # noqa: E402  (the emulated-ranks harness)

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
HP = dict(synth.HP_ENC)


def close(a, b, atol=1e-5):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all(np.abs(a - b) <= atol + 1e-4 * np.abs(b)))


def fail(msg):
    print("MISMATCH", msg)
    sys.exit(1)


t_end = time.time() + budget
count = dict(generic=0, phn=0, emph=0, chunked=0)
while time.time() < t_end:
    kind = rng.integers(0, 10)
    if kind < 4:                                                   # generic geometry front-end
        n_fft = int(rng.choice([256, 320, 400, 512, 800, 1024]))
        win = int(rng.choice([n_fft, n_fft, max(64, n_fft // 2), max(64, (n_fft * 3 // 4) // 2 * 2)]))
        hop = int(rng.choice([40, 64, 80, 100, 128, 160, 200]))
        kw = dict(HP)
        kw.update(n_fft=n_fft, win_length=win, hop_length=hop, n_mels=int(rng.choice([40, 64, 80, 128])),
                  n_mfcc=int(rng.choice([13, 20, 40])), window=str(rng.choice(["hann", "hamming"])),
                  calc_mfcc_derivate=bool(rng.integers(0, 2)))
        if kw["n_mfcc"] > kw["n_mels"]:
            kw["n_mfcc"] = kw["n_mels"]
        lens = [int(rng.integers(max(2 * hop, 32), 12000)) for _ in range(int(rng.integers(1, 5)))]
        wavs = [synth.utterance(int(rng.integers(0, 1 << 30)), n / 16000.0 + 0.01)[:n] for n in lens]
        got = al.calc_MFCC_input_batch(wavs, **kw)
        for y, g in zip(wavs, got):
            want = oracle.calc_MFCC_input(y, **kw)
            for name, a, b in zip(("mfcc", "mel", "pdb"), g, want):
                if not close(a, b):
                    fail(f"generic {name}: len {len(y)} geometry n_fft {n_fft} win {win} hop {hop} mels {kw['n_mels']} "
                         f"mfcc {kw['n_mfcc']} {kw['window']} max err {np.abs(a.astype(np.float64) - b).max():.3e}")
        count["generic"] += 1
    elif kind < 6:                                                 # frame labels
        hop, win = int(rng.choice([40, 80, 160])), int(rng.choice([200, 400, 800]))
        lens, phns = [], []
        for _ in range(int(rng.integers(1, 9))):
            n = int(rng.integers(win, 40000))
            cuts = np.sort(rng.choice(np.arange(1, n), size=int(rng.integers(1, 12)), replace=False))
            b = [0] + [int(c) for c in cuts] + [n]
            if rng.random() < 0.3:
                b[-1] = n - int(rng.integers(0, min(n - b[-2], 500)))          # labels may stop before the audio does
            phns.append([(b[i], b[i + 1], f"p{i % 5}") for i in range(len(b) - 1)])
            lens.append(n)
        conv = {f"p{i}": i for i in range(5)}
        got = al.calc_PHN_target_batch(lens, phns, conv, hop_length=hop, win_length=win)
        for n, p, g in zip(lens, phns, got):
            want = oracle.calc_PHN_target(np.zeros(n, np.float32), p, conv, hop, win)
            if not np.array_equal(g, np.asarray(want)):
                fail(f"phn: len {n} hop {hop} win {win} intervals {p}")
        count["phn"] += 1
    elif kind < 7:                                                 # emphasis filters (float32 and float64 input)
        n = int(rng.integers(1, 200000))
        c = float(rng.choice([0.97, 0.95, 0.5, 0.0, 0.999]))
        y = (0.1 * rng.standard_normal(n)).astype(rng.choice([np.float32, np.float64]))
        if c != 0.0:
            if not np.allclose(al.calc_preemphasis(y, c), oracle.calc_preemphasis(y, c), rtol=0, atol=1e-12):
                fail(f"preemphasis n {n} c {c} {y.dtype}")
            a, b = al.calc_inv_preemphasis(y, c), oracle.calc_inv_preemphasis(y, c)
            if not np.allclose(a, b, rtol=1e-9, atol=1e-10 * max(1.0, float(np.abs(b).max()))):
                fail(f"inv_preemphasis n {n} c {c} {y.dtype} max err {np.abs(a - b).max():.3e}")
        count["emph"] += 1
    else:                                                          # time-chunked Griffin-Lim, emulated ranks
        world = int(rng.integers(2, 5))
        geom_std = rng.random() < 0.8
        T = int(rng.integers(world * 340, world * 900))                 # every emulated rank owns at least one 112-frame block
        n_iter, k = int(rng.integers(2, 14)), int(rng.integers(1, 9))
        realse = float(rng.choice([1.0, 1.2]))
        try:
            if geom_std:
                tg.test_chunk_run_emulated_ranks_bit_identical(al, world, T, n_iter, k, realse)
            else:
                tg.test_chunk_run_emulated_ranks_bit_identical(al, 2, 321, n_iter, k, realse)      # the hop 40 / n_fft 800 geometry
        except AssertionError as exc:
            fail(f"chunked Griffin-Lim world {world} T {T} iterations {n_iter} k {k} realse {realse}: {str(exc)[:300]}")
        count["chunked"] += 1
print(f"soak2 ok in {budget:.0f} s: {count}; seed {seed}")
