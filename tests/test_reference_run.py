"""The oracle against the REFERENCE FILE ITSELF: tests/golden/reference_run_vectors.npz holds the outputs of
/root/reference/audio_lib.py executed unmodified under tests/golden/librosa_shim.py (make_reference_vectors.py).

* the oracle reproduces every vector (front-end to 2e-7 absolute = float32 rounding of the casts, waveforms to > 100 dB
  SNR, labels and emphasis filters exactly) - so the oracle's restatement of audio_lib.py's own code (gain, dtype chain,
  normalisation, deltas, clipping, Griffin-Lim loop with NumPy's random phase, realse) is pinned to the reference source;
* where /root/reference exists (this container, not the GPU box) the vectors are regenerated and must equal the file.

What stays restated are the librosa primitives inside the shim (pinned piecewise in tests/test_oracle_pins.py).
"""
import os

import numpy as np
import pytest

from oracle import audio_lib_oracle as oracle
from speech_cloner_b200 import synth
from tests.golden import make_reference_vectors as mr

G = np.load(mr.OUT)


def snr_db(got, want):
    return 10 * np.log10(np.sum(np.square(want, dtype=np.float64)) / max(np.sum((got - want) ** 2, dtype=np.float64), 1e-300))


def fe_case_input(name, seconds, ds_norm):
    return synth.utterance(int(G[f"{name}/seed"]), seconds, ds_norm=ds_norm)


def test_emphasis_filters_equal_the_reference_run():
    y = synth.utterance(4000, 0.25)
    assert (oracle.calc_preemphasis(y, 0.97) == G["preemph/out"]).all()
    assert (oracle.calc_inv_preemphasis(y, 0.97) == G["inv_preemph/out"]).all()


def test_phn_targets_equal_the_reference_run():
    for i, (n, phn_v) in enumerate(mr.phn_cases()):
        got = oracle.calc_PHN_target(np.zeros(n, np.float32), phn_v, mr.PHN_CONV, hop_length=80, win_length=400)
        assert (np.asarray(got) == G[f"phn{i}/target"]).all()


@pytest.mark.parametrize("case", mr.FE_CASES, ids=[c[0] for c in mr.FE_CASES])
def test_frontend_equals_the_reference_run(case):
    name, _, seconds, ds_norm, ov = case
    y = fe_case_input(name, seconds, ds_norm)
    got = oracle.calc_MFCC_input(y, **mr.fe_kwargs(ov))
    for g, key in zip(got, ("mfcc", "mel", "pdb")):
        want = G[f"{name}/{key}"]
        assert g.shape == want.shape and g.dtype == want.dtype
        np.testing.assert_allclose(g, want, rtol=0, atol=2e-7, err_msg=f"{name}/{key}")


def test_griffin_lim_equals_the_reference_run():
    """Same NumPy seed -> the reference's own np.random.rand phase (audio_lib.py:255)."""
    P_all = G["hp/pdb"]
    for name, seed, frames, n_iter, realse in mr.GL_CASES:
        np.random.seed(seed)
        got = oracle.from_power_to_wav(P_all[:frames], n_iter=n_iter, realse=realse, verbose=False, **mr.GL_KW)
        want = G[f"{name}/wav"]
        assert got.shape == want.shape and got.dtype == want.dtype
        assert snr_db(got, want) > 100.0, name
    F = np.sqrt(np.power(10.0, 0.1 * (P_all[:30].T / 0.01 - 80)))
    np.random.seed(6100)
    got = oracle.griffin_lim_alg(F, 400, 80, num_iters=5, verbose=False)
    assert got.dtype == G["gl_alg/wav"].dtype and snr_db(got, G["gl_alg/wav"]) > 100.0


@pytest.mark.skipif(not os.path.exists(mr.REFERENCE), reason="/root/reference is not on this machine")
def test_vectors_are_in_sync_with_the_reference_source():
    fresh = mr.compute(mr.load_reference())
    assert sorted(fresh) == sorted(G.files)
    for k, v in fresh.items():
        assert np.array_equal(np.asarray(v), G[k]), k
