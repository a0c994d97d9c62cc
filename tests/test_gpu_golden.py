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


# ---------------------------------------------------------------- the reference file itself (reference_run_vectors.npz)
from tests.golden import make_reference_vectors as mr  # noqa: E402

R = np.load(mr.OUT)


@pytest.mark.parametrize("case", mr.FE_CASES, ids=[c[0] for c in mr.FE_CASES])
def test_frontend_vs_the_reference_run(built_lib, case):
    """CUDA front-end vs the outputs of /root/reference/audio_lib.py executed under the librosa shim
    (make_reference_vectors.py), north-star tolerance 1e-4 relative / 1e-5 absolute."""
    from speech_cloner_b200 import audio_lib as al
    name, _, seconds, ds_norm, ov = case
    y = synth.utterance(int(R[f"{name}/seed"]), seconds, ds_norm=ds_norm)
    got = al.calc_MFCC_input(y, **mr.fe_kwargs(ov))
    for g, key in zip(got, ("mfcc", "mel", "pdb")):
        assert g.dtype == R[f"{name}/{key}"].dtype
        assert_close(g, R[f"{name}/{key}"], what=f"{name}/{key}")


def test_griffin_lim_vs_the_reference_run(built_lib):
    """Same NumPy seed: the product draws the reference's own ``np.pi * np.random.rand`` phase (audio_lib.py:255)."""
    from speech_cloner_b200 import audio_lib as al
    P_all = R["hp/pdb"]
    for name, seed, frames, n_iter, realse in mr.GL_CASES:
        np.random.seed(seed)
        got = al.from_power_to_wav(P_all[:frames], n_iter=n_iter, realse=realse, verbose=False, **mr.GL_KW)
        want = R[f"{name}/wav"]
        assert got.shape == want.shape and got.dtype == want.dtype
        assert snr_db(got, want) >= 40.0, name
    F = np.sqrt(np.power(10.0, 0.1 * (P_all[:30].T / 0.01 - 80)))
    np.random.seed(6100)
    got = al.griffin_lim_alg(F, 400, 80, num_iters=5, verbose=False)
    assert snr_db(got, R["gl_alg/wav"]) >= 40.0


def test_emphasis_and_labels_vs_the_reference_run(built_lib):
    from speech_cloner_b200 import audio_lib as al
    y = synth.utterance(4000, 0.25)
    np.testing.assert_allclose(al.calc_preemphasis(y, 0.97), R["preemph/out"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(al.calc_inv_preemphasis(y, 0.97), R["inv_preemph/out"], rtol=0, atol=1e-12)
    cases = mr.phn_cases()
    got = al.calc_PHN_target_batch([n for n, _ in cases], [p for _, p in cases], mr.PHN_CONV, hop_length=80, win_length=400)
    for i, g in enumerate(got):
        assert (g == R[f"phn{i}/target"]).all()
        n, phn_v = cases[i]
        assert (al.calc_PHN_target(np.zeros(n, np.float32), phn_v, mr.PHN_CONV, 80, 400) == R[f"phn{i}/target"]).all()
