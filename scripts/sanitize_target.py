"""Small target for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every kernel family of the library
on small ragged inputs.  Run through scripts/sanitize.sh on a B200.

Front-end: hp fast path (warp-specialised pass A, short-utterance pass A, pass B3, |y| sum), both FFT precisions, the
no-delta / no-clip variants, a generic geometry; Griffin-Lim: initial inverse STFT, one-tile and persistent iteration
kernels, rms deltas, the prologue / epilogue of from_power_to_wav, labels; emphasis filters; window gather.
"""
import sys

sys.path.insert(0, '.')
import numpy as np
import torch

from speech_cloner_b200 import audio_lib as al, dataset_cache as dc, synth

hp = dict(synth.HP_ENC)
wavs = synth.batch(3, 5, 0.5) + [synth.utterance(77, 0.05), synth.utterance(78, 0.28), synth.utterance(79, 1.3)]
for prec in ("fp64", "fp32"):
    feats = al.calc_MFCC_input_batch(wavs, fft_precision=prec, **hp)
feats = al.calc_MFCC_input_batch(wavs[:3], **{**hp, "calc_mfcc_derivate": False, "clip_output": False, "window": "hamming"})
gen = al.calc_MFCC_input_batch(wavs[:2])                                  # signature defaults: hop 40, 128 mels (generic kernels)
gen2 = al.calc_MFCC_input_batch(wavs[:2], **{**hp, "hop_length": 100, "win_length": 256, "n_fft": 512, "n_mels": 40, "n_mfcc": 13})
y = al.calc_preemphasis(wavs[0]); z = al.calc_inv_preemphasis(wavs[0])
phn = al.calc_PHN_target_batch([len(w) for w in wavs[:3]], [[(0, 3000, "a"), (3000, len(w), "b")] for w in wavs[:3]], {"a": 0, "b": 1}, 80, 400)

P = al.calc_MFCC_input(synth.utterance(5, 1.0), **hp)[2]
np.random.seed(0)
kw = dict(P_dB_norm_factor=0.01, pre_emphasis=0.97, hop_length=80, win_length=400, mean_abs_amp_norm=0.045, verbose=False)
w1 = al.from_power_to_wav(P[:70], n_iter=4, realse=1.2, **kw)
w2 = al.from_power_to_wav_batch([P[:150], P[:33], P[:90]], n_iter=3, **kw)
import io, contextlib
with contextlib.redirect_stdout(io.StringIO()):
    w3 = al.griffin_lim_alg(np.sqrt(np.power(10.0, 0.1 * (P[:40].T / 0.01 - 80))), 400, 80, num_iters=3, verbose=True)
w4 = al.from_power_to_wav((0.8 * np.random.rand(30, 401)).astype(np.float32), n_iter=2,
                          **{**kw, "hop_length": 40, "win_length": 800})                             # generic Griffin-Lim

lay = al.FrontendLayout([len(w) for w in wavs], 80)
plan = al._plan_from_kwargs(**hp)
dev = torch.zeros(lay.total_samples, dtype=torch.float32, device="cuda")
for w, o in zip(wavs, lay.sample_offsets):
    dev[o:o + len(w)] = torch.from_numpy(w).cuda()
m, l, p = al.frontend_device(plan, dev, lay)
cache = dc.DeviceSpecCache.from_device(m, l, p, lay)
np.random.seed(1)
nb = sum(1 for _ in dc.spec_window_sampler(cache, range(len(wavs)), 30, batch_size=3, prop_val=0.0, verbose=False))
torch.cuda.synchronize()
print("ok", nb, float(feats[0][2].mean()), float(np.abs(w1).mean()), al.launch_count())
