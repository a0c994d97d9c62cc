"""Small profiling target: 2 front-end steps (config-2 shape), a 2 048-window sampler gather out of their outputs and a
4-iteration Griffin-Lim (config-3 shape)."""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from speech_cloner_b200 import audio_lib as al, synth
prec = sys.argv[1] if len(sys.argv) > 1 else "fp64"
hp = dict(synth.HP_ENC)
base = synth.batch(2, 8, 4.0)
wavs = [base[i % 8] for i in range(256)]
kw = dict(sr=16000, n_fft=400, win_length=400, hop_length=80, n_mels=80, n_mfcc=40, window="hann", pre_emphasis=0.97,
          mfcc_normaleze_first_mfcc=True, mfcc_norm_factor=0.01, calc_mfcc_derivate=True, M_dB_norm_factor=0.01,
          P_dB_norm_factor=0.01, mean_abs_amp_norm=0.003, clip_output=True, fft_precision=prec)
plan = al.DspPlan(**kw)
lay = al.FrontendLayout([len(w) for w in wavs], 80)
dev = torch.zeros(lay.total_samples, dtype=torch.float32, device="cuda")
for w, o in zip(wavs, lay.sample_offsets):
    dev[o:o + len(w)] = torch.from_numpy(w).cuda()
for _ in range(2):
    out = al.frontend_device(plan, dev, lay)
torch.cuda.synchronize()
from speech_cloner_b200 import dataset_cache as dc
cache = dc.DeviceSpecCache.from_device(out[0], out[1], out[2], lay)
rs = np.random.RandomState(5)
u = rs.randint(0, 256, size=2048)
first = cache.frame_offsets[u] + rs.randint(0, np.asarray(cache.spec_len)[u] - 400)
win = cache.gather(dc.DeviceSpecCache.FEATURES, first, np.full(2048, 400, np.int32), 400)
torch.cuda.synchronize()
del win
glay = al._GlLayout([1000] * 64, 80)
gplan = al.DspPlan.get(n_fft=400, win_length=400, hop_length=80)
amp = torch.rand((glay.frame_offsets[-1], 201), device="cuda") * 0.1
ph = torch.rand((glay.frame_offsets[-1], 201), device="cuda") * 3.14159
w = al.griffin_lim_device(gplan, amp, ph, glay, 4)
torch.cuda.synchronize()
print("ok", float(out[2].mean()), float(w.abs().mean()))
