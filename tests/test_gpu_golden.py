"""CUDA path vs the committed known-answer vectors (tests/golden/oracle_vectors.npz), through the C ABI."""
import os

import numpy as np
import pytest

from speech_cloner_b200 import synth
from tests.golden import make_golden as mg
from tests.util import assert_close, snr_db

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.npz"))


def test_frontend_golden(built_lib):
    from speech_cloner_b200 import audio_lib as al
    for name, seed, seconds, ds_norm, ov in mg.FE_CASES:
        kw = dict(synth.HP_ENC); kw.update(ov)
        got = al.calc_MFCC_input(mg.fe_inputs(seed, seconds, ds_norm), **kw)
        for g, key in zip(got, ("mfcc", "mel", "pdb")):
            assert_close(g, G[f"{name}/{key}"], what=f"{name}/{key}")


def test_griffin_lim_golden(built_lib):
    from speech_cloner_b200 import audio_lib as al
    for name, seed, frames, n_iter, realse in mg.GL_CASES:
        P, ph = mg.gl_inputs(seed, frames)
        got = al.from_power_to_wav(P, n_iter=n_iter, realse=realse, verbose=False, phase0=ph, **mg.GL_KW)
        assert snr_db(got, G[f"{name}/wav"]) >= 40.0
