#!/usr/bin/env python
"""e2e (host buffers, pinned H2D + D2H inside the timed region) against the chunking of FrontendPipeline, and the raw
pinned D2H / H2D bandwidth of this box for the same byte counts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_cloner_b200 import audio_lib as al, synth

hp = dict(synth.HP_ENC)
kw = dict(sr=16000, n_fft=400, win_length=400, hop_length=80, n_mels=80, n_mfcc=40, window="hann", pre_emphasis=0.97,
          mfcc_normaleze_first_mfcc=True, mfcc_norm_factor=0.01, calc_mfcc_derivate=True, M_dB_norm_factor=0.01,
          P_dB_norm_factor=0.01, mean_abs_amp_norm=0.003, clip_output=True)
base = synth.batch(2, 16, 4.0)
wavs = [base[i % 16] for i in range(256)]
# raw copies
h = torch.empty(297209856 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(h, device="cuda")
hi = torch.empty(65536000 // 4, dtype=torch.float32).pin_memory()
di = torch.empty_like(hi, device="cuda")
for _ in range(3):
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(10):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 10
print(f"raw D2H 297 MB: {dt*1e3:.3f} ms = {297.2/dt/1e3:.1f} GB/s")
s2 = torch.cuda.Stream()
t = time.perf_counter()
for _ in range(10):
    h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        di.copy_(hi, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 10
print(f"raw D2H 297 MB with concurrent H2D 65 MB: {dt*1e3:.3f} ms")
for n_chunks, n_streams, ramp in [(8, 3, False), (8, 3, True), (6, 3, True), (12, 3, True), (8, 4, True), (8, 2, True)]:
    pipe = al.FrontendPipeline([len(w) for w in wavs], n_chunks=n_chunks, n_streams=n_streams, ramp=ramp, **kw)
    pipe.load(wavs)
    for _ in range(3):
        pipe.run()
    t = time.perf_counter()
    for _ in range(10):
        pipe.run()
    dt = (time.perf_counter() - t) / 10
    print(f"chunks {n_chunks:3d} ({len(pipe.chunks)} actual) streams {n_streams} ramp {ramp}: {dt*1e3:.3f} ms  {1024/dt:.0f} audio-s/s")
    del pipe
