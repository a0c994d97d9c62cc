"""GPU parity of the dataset-cache path: sc_phn_target_batch against the oracle's literal
calc_PHN_target, and the batched cache builder against the reference loop (oracle calc_MFCC_input per utterance).
"""
import numpy as np
import pytest

from oracle import audio_lib_oracle as oracle
from speech_cloner_b200 import synth
from tests.test_dataset_cache import CFG, _random_case
from tests.util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def al(built_lib):
    from speech_cloner_b200 import audio_lib
    return audio_lib


def test_phn_target_batch_is_bit_exact(al):
    rng = np.random.default_rng(11)
    conv = {f"p{i}": i for i in range(7)}
    cases = [_random_case(rng) for _ in range(64)]
    hop, win = 80, 400
    lens = [len(c[0]) for c in cases]
    phns = [c[1] for c in cases]
    got = al.calc_PHN_target_batch(lens, phns, conv, hop_length=hop, win_length=win)
    idx = al.calc_PHN_target_batch(lens, phns, conv, hop_length=hop, win_length=win, return_index=True)
    for (y, phn, _, _), g, i in zip(cases, got, idx):
        want = oracle.calc_PHN_target(y, phn, conv, hop, win)
        assert g.dtype == np.int32 and g.shape == want.shape and (g == want).all()
        assert i.min() >= 0 and i.max() < len(phn)


def test_phn_target_batch_rejects_unsorted_ends(al):
    y, phn, hop, win = np.zeros(4000, np.float32), [(0, 3000, "a"), (3000, 2000, "b"), (2000, 4000, "c")], 80, 400
    with pytest.raises(ValueError):
        al.calc_PHN_target_batch([len(y)], [phn], {"a": 0, "b": 1, "c": 2}, hop, win)


def test_cache_builder_equals_the_reference_loop(al, tmp_path):
    from speech_cloner_b200 import dataset_cache as dc
    rng = np.random.default_rng(5)
    wavs = synth.batch(7, 6, 1.0, ds_norm=(0.0, 10.0)) + [synth.utterance(7100, 0.37), synth.utterance(7101, 2.3)]
    conv = {f"p{i}": np.eye(7, dtype=np.int32)[i] for i in range(7)}
    phn_vs = []
    for w in wavs:
        cuts = np.sort(rng.integers(0, len(w), size=5))
        b = [0] + list(cuts) + [len(w)]
        phn_vs.append([(int(b[i]), int(b[i + 1]), f"p{i % 7}") for i in range(len(b) - 1)])
    cfg = dict(CFG, n_fft=400)
    path = str(tmp_path / dc.spec_cache_name(cfg, "TIMIT"))
    used = dc.build_spec_cache({"wav": wavs, "phn_v": phn_vs}, cfg, path, phn_conv_d=conv, fmt="npz",
                               max_batch_samples=40000)                       # several batches
    assert used == "npz"
    assert dc.build_spec_cache({"wav": wavs, "phn_v": phn_vs}, cfg, path, phn_conv_d=conv) is None   # exists: untouched
    cache = dc.open_cache(path)
    kw = dc._frontend_kwargs(cfg)
    for i, (w, phn) in enumerate(zip(wavs, phn_vs)):
        want = oracle.calc_MFCC_input(w, **kw)
        for name, ref in zip(("mfcc", "mel_dB", "power_dB"), want):
            got = cache[name][str(i)][...]
            assert got.dtype == np.float32
            assert_close(got, ref, what=f"sample {i} {name}")
        want_phn = oracle.calc_PHN_target(w, phn, conv, cfg["hop_length"], cfg["win_length"])
        assert (cache["phn"][str(i)][...] == want_phn).all()
    cache.close()


def test_cache_builder_without_labels(al, tmp_path):
    """TARGET_spk_reader.py:132-182 stores no phn group."""
    from speech_cloner_b200 import dataset_cache as dc
    wavs = synth.batch(9, 3, 0.5)
    cfg = dict(CFG, n_fft=400)
    path = str(tmp_path / dc.spec_cache_name(cfg, "TARGET"))
    dc.build_spec_cache({"wav": wavs}, cfg, path, fmt="npz")
    cache = dc.open_cache(path)
    assert len(cache["mfcc"]) == 3 and len(cache["phn"]) == 0
    cache.close()
