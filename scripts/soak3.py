"""Third randomised soak: the host pipeline (FrontendPipeline, the e2e path) against the one-call batch for random
chunkings, device samplers against the oracle's literal loops, and the conversion glue on CUDA tensors against the oracle.

    python scripts/soak3.py [seconds] [seed]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import audio_lib_oracle as oracle                          # noqa: E402
from speech_cloner_b200 import audio_lib as al, conversion as cv, dataset_cache as dc, synth   # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
HP = dict(synth.HP_ENC)
PLAN_KW = dict(sr=16000, n_fft=400, win_length=400, hop_length=80, n_mels=80, n_mfcc=40, window="hann", pre_emphasis=0.97,
               mfcc_normaleze_first_mfcc=True, mfcc_norm_factor=0.01, calc_mfcc_derivate=True, M_dB_norm_factor=0.01,
               P_dB_norm_factor=0.01, mean_abs_amp_norm=0.003, clip_output=True)


def fail(msg):
    print("MISMATCH", msg)
    sys.exit(1)


class DictCache(dict):
    pass


t_end = time.time() + budget
count = dict(pipeline=0, samplers=0, conversion=0)
while time.time() < t_end:
    kind = rng.integers(0, 3)
    if kind == 0:
        lens = [int(rng.choice([rng.integers(300, 3000), rng.integers(3000, 40000), 4496 + rng.integers(-2, 3), 15999, 8001]))
                for _ in range(int(rng.integers(1, 20)))]
        wavs = [synth.utterance(int(rng.integers(0, 1 << 30)), n / 16000.0 + 0.01)[:n] for n in lens]
        n_chunks, n_streams = int(rng.integers(1, 12)), int(rng.integers(1, 5))
        pipe = al.FrontendPipeline(lens, n_chunks=n_chunks, n_streams=n_streams, ramp=bool(rng.integers(0, 2)), **PLAN_KW)
        pipe.load(wavs)
        pipe.run()
        got = [tuple(a.copy() for a in v) for v in pipe.views()]
        want = al.calc_MFCC_input_batch(wavs, **HP)
        for i, (g, w) in enumerate(zip(got, want)):
            for a, b in zip(g, w):
                if not np.array_equal(a, b):
                    fail(f"pipeline != batch: utterance {i} of lens {lens}, chunks {n_chunks} streams {n_streams}")
        count["pipeline"] += 1
    elif kind == 1:
        lens = [int(rng.integers(3, 120)) for _ in range(int(rng.integers(4, 30)))]
        n_t = int(rng.integers(5, 60))
        widths = dict(mfcc=int(rng.choice([6, 80, 3])), mel_dB=int(rng.choice([5, 80, 1])), power_dB=int(rng.choice([7, 201, 2])))
        host = {g: {str(i): rng.random((n, w)).astype(np.float32) for i, n in enumerate(lens)} for g, w in widths.items()}
        host["phn"] = {str(i): rng.integers(0, 61, size=n).astype(np.int32) for i, n in enumerate(lens)}
        keys = [str(i) for i in range(len(lens))]
        dev = dc.DeviceSpecCache.from_arrays({g: [host[g][k] for k in keys] for g in host}, keys)
        ids = np.sort(rng.choice(len(lens), size=int(rng.integers(2, len(lens) + 1)), replace=False))
        kw = dict(batch_size=int(rng.integers(1, 6)), n_epochs=int(rng.integers(1, 3)), randomize_samples=bool(rng.integers(0, 2)),
                  sample_trn=bool(rng.integers(0, 2)), prop_val=float(rng.choice([0.0, 0.3, 0.5])), random_seed=int(rng.integers(0, 99)),
                  yield_idxs=True)
        if kw["prop_val"] > 0 and int(kw["prop_val"] * len(ids)) == 0:
            kw["prop_val"] = 0.0                                       # idx_v[:-0] is empty in the reference: not a useful case
        np.random.seed(3)
        want = list(oracle.spec_window_sampler(host, ids, n_t, **kw))
        st = np.random.get_state()[1].copy()
        np.random.seed(3)
        got = list(dc.spec_window_sampler(dev, ids, n_t, verbose=False, **kw))
        if len(got) != len(want) or not np.array_equal(np.random.get_state()[1], st):
            fail(f"spec_window_sampler: {len(got)} vs {len(want)} batches or random stream differs; lens {lens} n_t {n_t} {kw}")
        for g, w in zip(got, want):
            if not np.array_equal(g[3], w[3]) or any(not np.array_equal(a.cpu().numpy(), b.astype(np.float32)) for a, b in zip(g[:3], w[:3])):
                fail(f"spec_window_sampler batch content; lens {lens} n_t {n_t} {kw}")
        np.random.seed(4)
        want = list(oracle.window_sampler(host, ids, n_t, batch_size=kw["batch_size"], n_epochs=kw["n_epochs"], yield_idxs=True))
        np.random.seed(4)
        got = list(dc.window_sampler(dev, ids, n_t, batch_size=kw["batch_size"], n_epochs=kw["n_epochs"], yield_idxs=True))
        if len(got) != len(want):
            fail("window_sampler batch count")
        for (x, y, i1), (wx, wy, i2) in zip(got, want):
            if not (np.array_equal(i1, i2) and np.array_equal(x.cpu().numpy(), wx) and np.array_equal(y.cpu().numpy(), wy)):
                fail(f"window_sampler content; lens {lens} n_t {n_t}")
        count["samplers"] += 1
    else:
        n_t = int(rng.choice([8, 20, 40]))
        cfg = dict(hop_length=80, n_timesteps=n_t, sample_rate=16000)
        T = int(rng.integers(n_t + 1, 40 * n_t))
        mf, ml, sf = rng.random((T, 6)).astype(np.float32), rng.random((T, 5)).astype(np.float32), rng.random((T, 7)).astype(np.float32)
        t_s, t_e = 0, int(rng.integers(1, 4))
        try:
            w = oracle.window_batches(mf, ml, sf, cfg, t_s=t_s, t_e=t_e)
        except Exception:
            w = None
        try:
            g = cv.window_batches(torch.from_numpy(mf).cuda(), torch.from_numpy(ml).cuda(), torch.from_numpy(sf).cuda(), cfg, t_s=t_s, t_e=t_e)
        except Exception:
            g = None
        if (w is None) != (g is None):
            fail(f"window_batches error parity T {T} n_t {n_t} t_e {t_e}")
        if w is not None:
            pairs = [(g["mfcc_input0"], w[0]), (g["mfcc_input1"], w[1]), (g["mel_true"], w[2]), (g["stft_true"], w[3])]
            for a, b in pairs:
                if (a is None) != (b is None) or (a is not None and not np.array_equal(a.cpu().numpy(), np.asarray(b, dtype=np.float32))):
                    fail(f"window_batches content T {T} n_t {n_t} t_e {t_e}")
            if (g["n_s"], g["n_e"]) != (w[4], w[5]):
                fail("window_batches bounds")
            p0 = rng.random(tuple(w[0].shape[:2]) + (4,))
            if w[1] is not None:
                p1 = rng.random(tuple(w[1].shape[:2]) + (4,))
                if not np.array_equal(cv.compound(torch.from_numpy(p0).cuda(), torch.from_numpy(p1).cuda()).cpu().numpy(), oracle.compound(p0, p1)):
                    fail(f"compound N {p0.shape[0]} T {n_t}")
        count["conversion"] += 1
print(f"soak3 ok in {budget:.0f} s: {count}; seed {seed}")
