"""Profiling target: a 6-iteration Griffin-Lim of the config-3 shape (64 x 1000 frames)."""
import sys; sys.path.insert(0, '.')
import torch
from speech_cloner_b200 import audio_lib as al
glay = al._GlLayout([1000] * 64, 80)
gplan = al.DspPlan.get(n_fft=400, win_length=400, hop_length=80)
amp = torch.rand((glay.frame_offsets[-1], 201), device="cuda") * 0.1
ph = torch.rand((glay.frame_offsets[-1], 201), device="cuda") * 3.14159
w = al.griffin_lim_device(gplan, amp, ph, glay, 6)
torch.cuda.synchronize()
print("ok", float(w.abs().mean()))
