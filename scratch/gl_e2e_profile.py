"""cProfile of the host side of from_power_to_wav_batch (config-3 shape, host arrays)."""
import cProfile, pstats, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from speech_cloner_b200 import audio_lib as al, synth
hp = dict(synth.HP_ENC)
wavs = synth.batch(3, 64, 5.0)
feats = al.calc_MFCC_input_batch(wavs, **hp)
Ps = [np.ascontiguousarray(f[2][:1000]) for f in feats]
phs = []
for i in range(64):
    np.random.seed(3000 + i); phs.append(np.pi * np.random.rand(201, 1000))
kw = dict(P_dB_norm_factor=0.01, pre_emphasis=0.97, hop_length=80, win_length=400, mean_abs_amp_norm=0.045, n_iter=200, realse=1.0, verbose=False)
for _ in range(2):
    al.from_power_to_wav_batch(Ps, phase0s=phs, **kw)
torch.cuda.synchronize()
t = time.perf_counter(); al.from_power_to_wav_batch(Ps, phase0s=phs, **kw); print("wall ms", (time.perf_counter() - t) * 1e3)
pr = cProfile.Profile(); pr.enable(); al.from_power_to_wav_batch(Ps, phase0s=phs, **kw); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
