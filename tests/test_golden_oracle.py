"""The oracle must keep reproducing the frozen known-answer vectors (tests/golden/, made by make_golden.py)."""
import os

import numpy as np

from oracle import audio_lib_oracle as oracle
from speech_cloner_b200 import synth
from tests.golden import make_golden as mg

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.npz"))


def test_frontend_vectors():
    for name, seed, seconds, ds_norm, ov in mg.FE_CASES:
        kw = dict(synth.HP_ENC); kw.update(ov)
        got = oracle.calc_MFCC_input(mg.fe_inputs(seed, seconds, ds_norm), **kw)
        for g, key in zip(got, ("mfcc", "mel", "pdb")):
            np.testing.assert_allclose(g, G[f"{name}/{key}"], rtol=0, atol=2e-7, err_msg=f"{name}/{key}")


def test_griffin_lim_vectors():
    for name, seed, frames, n_iter, realse in mg.GL_CASES:
        P, ph = mg.gl_inputs(seed, frames)
        got = oracle.from_power_to_wav(P, n_iter=n_iter, realse=realse, verbose=False, phase0=ph, **mg.GL_KW)
        want = G[f"{name}/wav"]
        assert 10 * np.log10(np.sum(want ** 2) / max(np.sum((got - want) ** 2), 1e-300)) > 100.0
