#!/usr/bin/env python
"""A/B timing of the front-end step (configs[1] shape) per kernel, for the library named by
SPEECHDSP_LIB (default: the in-tree build) and the env knobs the library reads (SC_FE_WS, ...).

    python scripts/ab_bench.py [fp64|fp32] [steps]

Prints one line: step ms back to back + the in-library CUDA-event split (gain | pass A | pass B),
and checks the result against the first call's output of the same process (determinism)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from speech_cloner_b200 import audio_lib as al, synth

prec = sys.argv[1] if len(sys.argv) > 1 else "fp64"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
hp = dict(synth.HP_ENC)
base = synth.batch(2, 16, 4.0)
wavs = [base[i % 16] for i in range(256)]
kw = dict(sr=16000, n_fft=400, win_length=400, hop_length=80, n_mels=80, n_mfcc=40, window="hann", pre_emphasis=0.97,
          mfcc_normaleze_first_mfcc=True, mfcc_norm_factor=0.01, calc_mfcc_derivate=True, M_dB_norm_factor=0.01,
          P_dB_norm_factor=0.01, mean_abs_amp_norm=0.003, clip_output=True, fft_precision=prec)
plan = al.DspPlan(**kw)
lay = al.FrontendLayout([len(w) for w in wavs], 80)
dev = torch.zeros(lay.total_samples, dtype=torch.float32, device="cuda")
for w, o in zip(wavs, lay.sample_offsets):
    dev[o:o + len(w)] = torch.from_numpy(w).cuda()
out = al.frontend_device(plan, dev, lay)
ref = [o.clone() for o in out]
for _ in range(5):
    al.frontend_device(plan, dev, lay, out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    al.frontend_device(plan, dev, lay, out)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
plan.profile(True)
prof = np.zeros(3)
for _ in range(20):
    al.frontend_device(plan, dev, lay, out)
    prof += np.array(plan.profile_read()[:3])
plan.profile(False)
prof /= 20
same = all(bool(torch.equal(x, y)) for x, y in zip(out, ref))
tag = os.environ.get("SPEECHDSP_LIB", "in-tree") + " " + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("SC_"))
print(f"[{tag}] {prec} step {ms:.4f} ms | gain {prof[0]:.4f} passA {prof[1]:.4f} passB {prof[2]:.4f} | "
      f"frac {1764 * 205056 / (ms * 1e-3) / 1e9 / 6553.9:.4f} | deterministic {same} | mean {float(out[2].mean()):.6f}")
