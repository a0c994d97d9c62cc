"""Generate the known-answer fixtures in this directory from the CPU oracle.

    python tests/golden/make_golden.py

These vectors freeze the ORACLE (oracle/audio_lib_oracle.py): /root/reference ships no tests or vectors of its own.  The
oracle is pinned to the reference file's own run by make_reference_vectors.py / tests/test_reference_run.py and, for the
librosa primitives, piecewise against torch.stft / transformers / scipy / torchaudio in tests/test_oracle_pins.py.
They guard both the oracle and the CUDA path against drift.
Inputs are regenerated from seeds by speech_cloner_b200.synth, so only outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import audio_lib_oracle as oracle  # noqa: E402
from speech_cloner_b200 import synth  # noqa: E402

FE_CASES = [  # (name, seed, seconds, ds_norm, overrides)
    ("timit_hp", 1000, 0.6, (0.0, 10.0), {}),
    ("arctic_hp", 2000, 0.5, (0.0, 1.0), {}),
    ("nodelta_hamming", 2001, 0.4, (0.0, 1.0), {"calc_mfcc_derivate": False, "window": "hamming"}),
]
GL_CASES = [  # (name, seed, frames, n_iter, realse)
    ("gl_r10", 3000, 60, 12, 1.0),
    ("gl_r12", 3001, 48, 8, 1.2),
]
GL_KW = dict(P_dB_norm_factor=0.01, pre_emphasis=0.97, hop_length=80, win_length=400, mean_abs_amp_norm=0.045)


def fe_inputs(seed, seconds, ds_norm):
    return synth.utterance(seed, seconds, ds_norm=ds_norm)


def gl_inputs(seed, frames):
    P = oracle.calc_MFCC_input(synth.utterance(seed, 1.0), **synth.HP_ENC)[2][:frames]
    np.random.seed(seed)
    return P, np.pi * np.random.rand(201, frames)


def main():
    out = {}
    for name, seed, seconds, ds_norm, ov in FE_CASES:
        kw = dict(synth.HP_ENC); kw.update(ov)
        mfcc, mel, pdb = oracle.calc_MFCC_input(fe_inputs(seed, seconds, ds_norm), **kw)
        out[f"{name}/mfcc"], out[f"{name}/mel"], out[f"{name}/pdb"] = mfcc, mel, pdb
    for name, seed, frames, n_iter, realse in GL_CASES:
        P, ph = gl_inputs(seed, frames)
        out[f"{name}/wav"] = oracle.from_power_to_wav(P, n_iter=n_iter, realse=realse, verbose=False, phase0=ph, **GL_KW)
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
