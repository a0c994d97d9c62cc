"""The oracle's and the product's caller-side code against the REFERENCE'S OWN caller code
(tests/golden/reference_caller_vectors.npz: functions cut out of /root/reference/test.py, sound_ds.py and
TIMIT_reader.py with ``ast`` and executed unmodified, see make_reference_caller_vectors.py).  No GPU needed here; the
device samplers and the GPU cache builder are checked against the same vectors in tests/test_gpu_dataset_cache.py.
"""
import os

import numpy as np
import pytest

from oracle import audio_lib_oracle as oracle
from tests.golden import make_reference_caller_vectors as mc

G = np.load(mc.OUT)


def batches(prefix, names):
    out, b = [], 0
    while f"{prefix}/{b}/idxs" in G.files:
        out.append(tuple(G[f"{prefix}/{b}/{n}"] for n in names))
        b += 1
    return out


def test_compound_equals_the_reference_function():
    from speech_cloner_b200 import conversion as cv
    for i, (y0, y1) in enumerate(mc.compound_inputs()):
        want = G[f"compound{i}/out"]
        assert np.array_equal(oracle.compound(y0, y1), want)
        assert np.array_equal(cv.compound(y0, y1), want)


def test_oracle_cache_loop_equals_the_reference_loop():
    """calc_MFCC_input + calc_PHN_target per utterance, as TIMIT.create_phn_mfcc_cache stores them."""
    wavs, phn_vs = mc.cache_inputs()
    from speech_cloner_b200 import dataset_cache as dc
    kw = dc._frontend_kwargs(mc.CFG)
    for i, (y, phn_v) in enumerate(zip(wavs, phn_vs)):
        feats = oracle.calc_MFCC_input(y, **kw)
        for g, a in zip(("mfcc", "mel_dB", "power_dB"), feats):
            assert np.array_equal(a, G[f"cache/{g}/{i}"]), (g, i)
        phn = oracle.calc_PHN_target(y, phn_v, mc.mr.PHN_CONV, hop_length=80, win_length=400)
        assert np.array_equal(np.asarray(phn), G[f"cache/phn/{i}"])


def test_oracle_spec_window_sampler_equals_the_reference_method():
    cache = mc.sampler_cache("mem://oracle")
    ids = np.arange(len(mc.SAMPLER_LENS))[mc.sampler_filter()]
    for r, kw in enumerate(mc.SPEC_RUNS):
        want = batches(f"spec{r}", ("mfcc", "mel_dB", "power_dB", "idxs"))
        got = list(oracle.spec_window_sampler(cache, ids, mc.N_TIMESTEPS, random_seed=mc.RANDOM_SEED, yield_idxs=True, **kw))
        assert len(got) == len(want) > 0
        for g, w in zip(got, want):
            for a, b in zip(g, w):
                assert a.dtype == b.dtype and np.array_equal(a, b)
        assert np.array_equal(np.random.get_state()[1][:8], G[f"spec{r}/rng_after"])


def test_oracle_window_sampler_equals_the_reference_method():
    cache = mc.sampler_cache("mem://oracle")
    ids = np.arange(len(mc.SAMPLER_LENS))[mc.sampler_filter()]
    for r, kw in enumerate(mc.WIN_RUNS):
        want = batches(f"win{r}", ("x", "y", "idxs"))
        np.random.seed(100 + r)
        got = list(oracle.window_sampler(cache, ids, mc.N_TIMESTEPS, yield_idxs=True, **kw))
        assert len(got) == len(want) > 0
        for g, w in zip(got, want):
            for a, b in zip(g, w):
                assert a.dtype == b.dtype and np.array_equal(a, b)


@pytest.mark.skipif(not os.path.isdir(mc.REF_DIR), reason="/root/reference is not on this machine")
def test_vectors_are_in_sync_with_the_reference_source():
    fresh = mc.compute()
    assert sorted(fresh) == sorted(G.files)
    for k, v in fresh.items():
        assert np.array_equal(np.asarray(v), G[k]), k
