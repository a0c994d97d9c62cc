"""Host logic of the batched dataset-cache builder and of calc_PHN_target: no GPU needed.

* cache file names equal the reference's md5 recipe (TIMIT_reader.py:92-111, ARCTIC_reader.py:62-79)
* the vectorised host calc_PHN_target equals the oracle's literal transcription of audio_lib.py:51-85,
  including unsorted interval ends (cursor semantics)
* the npz cache layout round-trips under the reference's group / key names
"""
import hashlib

import numpy as np
import pytest

from oracle import audio_lib_oracle as oracle

CFG = dict(use_all_phonemes=True, sample_rate=16000, pre_emphasis=0.97, hop_length=80, win_length=400, n_mels=80,
           n_mfcc=40, n_fft=None, window="hann", mfcc_normaleze_first_mfcc=True, mfcc_norm_factor=0.01,
           calc_mfcc_derivate=True, M_dB_norm_factor=0.01, P_dB_norm_factor=0.01, mean_abs_amp_norm=0.003,
           clip_output=True, phn_mfcc_cache_name="phn_mfcc_cache.h5py", spec_cache_name="spec_cache.h5py")


def test_cache_names_follow_the_reference_md5_recipe():
    from speech_cloner_b200 import dataset_cache as dc
    keys = ("sample_rate", "pre_emphasis", "hop_length", "win_length", "n_mels", "n_mfcc", "n_fft", "window",
            "mfcc_normaleze_first_mfcc", "mfcc_norm_factor", "calc_mfcc_derivate", "M_dB_norm_factor",
            "P_dB_norm_factor", "mean_abs_amp_norm", "clip_output")
    h_spec = hashlib.md5("_".join(str(CFG[k]) for k in keys).encode()).hexdigest()
    h_timit = hashlib.md5("_".join(str(CFG[k]) for k in ("use_all_phonemes",) + keys).encode()).hexdigest()
    assert dc.spec_cache_name(CFG, "ARCTIC") == f"spec_cache_{h_spec}.h5py"
    assert dc.spec_cache_name(CFG, "TARGET") == f"spec_cache_{h_spec}.h5py"
    assert dc.spec_cache_name(CFG, "TIMIT") == f"phn_mfcc_cache_{h_timit}.h5py"
    other = dict(CFG, hop_length=40)
    assert dc.spec_cache_name(other, "ARCTIC") != dc.spec_cache_name(CFG, "ARCTIC")


def _random_case(rng, sorted_ends=True):
    L = int(rng.integers(100, 50000))
    hop = int(rng.choice([40, 80, 160]))
    win = int(rng.choice([200, 400, 800]))
    k = int(rng.integers(1, 30))
    cuts = np.sort(rng.integers(0, L + 500, size=k - 1)) if k > 1 else np.array([], dtype=int)
    b = [0] + list(cuts) + [max([L] + list(cuts)) + int(rng.integers(0, 300))]      # intervals may overrun the audio
    phn = [(int(b[i]), int(b[i + 1]), f"p{i % 7}") for i in range(len(b) - 1)]
    if not sorted_ends and len(phn) > 2:
        i = int(rng.integers(0, len(phn) - 1))
        phn[i], phn[i + 1] = phn[i + 1], phn[i]
    return np.zeros(L, np.float32), phn, hop, win


@pytest.mark.parametrize("sorted_ends", [True, False])
def test_host_phn_target_equals_the_literal_loop(sorted_ends):
    from speech_cloner_b200 import audio_lib as al
    rng = np.random.default_rng(7 if sorted_ends else 8)
    conv = {f"p{i}": i for i in range(7)}
    for _ in range(150):
        y, phn, hop, win = _random_case(rng, sorted_ends)
        want = oracle.calc_PHN_target(y, phn, conv, hop, win)
        got = al.calc_PHN_target(y, phn, conv, hop, win)
        assert got.dtype == np.int32 and got.shape == want.shape
        assert (got == want).all()


def test_phn_target_one_hot_labels():
    """phn_conv_d maps to one-hot vectors in the readers (TIMIT_reader.py:121-127): rows, not scalars."""
    from speech_cloner_b200 import audio_lib as al
    conv = {f"p{i}": np.eye(7, dtype=np.int32)[i] for i in range(7)}
    y, phn, hop, win = _random_case(np.random.default_rng(3))
    want = oracle.calc_PHN_target(y, phn, conv, hop, win)
    got = al.calc_PHN_target(y, phn, conv, hop, win)
    assert got.shape == want.shape == (len(y) // hop + 1, 7)
    assert (got == want).all()


def test_npz_cache_layout_round_trip(tmp_path):
    from speech_cloner_b200 import dataset_cache as dc
    path = str(tmp_path / dc.spec_cache_name(CFG, "ARCTIC"))
    w, used = dc._open_writer(path, "npz")
    assert used == "npz"
    with w as out:
        g = {k: out.create_group(k) for k in ("mfcc", "mel_dB", "power_dB", "phn")}
        for i in range(3):
            g["mfcc"].create_dataset(str(i), data=np.full((4 + i, 80), i, np.float32))
            g["mel_dB"].create_dataset(str(i), data=np.full((4 + i, 80), -i, np.float32))
            g["power_dB"].create_dataset(str(i), data=np.zeros((4 + i, 201), np.float32))
            g["phn"].create_dataset(str(i), data=np.arange(4 + i, dtype=np.int32))
    cache = dc.open_cache(path)
    assert sorted(cache["mfcc"].keys()) == ["0", "1", "2"] and len(cache["phn"]) == 3
    assert cache["mfcc"]["2"].shape == (6, 80) and cache["mfcc"]["2"][0, 0] == 2.0
    assert cache["power_dB"]["1"].shape == (5, 201)
    assert (cache["phn"]["0"] == np.arange(4)).all()
    cache.close()


def test_batches_respect_both_limits():
    from speech_cloner_b200 import dataset_cache as dc
    lens = [10, 20, 30, 40, 50, 5, 5, 5]
    got = [list(r) for r in dc._batches(lens, max_samples=60, max_utts=3)]
    assert [i for b in got for i in b] == list(range(len(lens)))
    for b in got:
        assert len(b) <= 3 and (sum(lens[i] for i in b) <= 60 or len(b) == 1)


def _toy_cache(tmp_path, lens):
    from speech_cloner_b200 import dataset_cache as dc
    path = str(tmp_path / "toy_cache.h5py")
    rng = np.random.default_rng(1)
    w, _ = dc._open_writer(path, "npz")
    with w as out:
        g = {k: out.create_group(k) for k in ("mfcc", "mel_dB", "power_dB", "phn")}
        for i, n in enumerate(lens):
            g["mfcc"].create_dataset(str(i), data=rng.random((n, 6)).astype(np.float32))
            g["mel_dB"].create_dataset(str(i), data=rng.random((n, 5)).astype(np.float32))
            g["power_dB"].create_dataset(str(i), data=rng.random((n, 7)).astype(np.float32))
            g["phn"].create_dataset(str(i), data=rng.integers(0, 9, size=n).astype(np.int32))
    return dc.open_cache(path)




# ------------------------------------------------------------------------------------------- window samplers (oracle)
def test_oracle_window_sampler_follows_the_reference_rng_order(tmp_path):
    """TIMIT_reader.py:474-523: shuffle once per epoch, skip short utterances, one randint per kept utterance.
    The oracle's transcription against an independent replay of the same random stream."""
    lens = [50, 12, 33, 8, 70, 41, 10, 29, 64, 55]
    cache = _toy_cache(tmp_path, lens)
    ids = [0, 1, 2, 3, 4, 5, 7, 8, 9]
    np.random.seed(42)
    got = list(oracle.window_sampler(cache, ids, n_timesteps=20, batch_size=3, n_epochs=2, yield_idxs=True))
    np.random.seed(42)
    order = [str(i) for i in ids]
    want = []
    for _ in range(2):
        np.random.shuffle(order)
        for s in order:
            n = lens[int(s)]
            if n <= 20:
                continue
            i_s = np.random.randint(0, n - 20)
            want.append((i_s, i_s + 20, int(s)))
    flat = [tuple(r) for b in got for r in b[2]]
    assert flat == want[: len(flat)] and len(flat) == 3 * (len(want) // 3)
    x, y, idx = got[0]
    assert x.shape == (3, 20, 6) and y.shape == (3, 20)
    i_s, i_e, s = idx[1]
    assert (x[1] == cache["mfcc"][str(s)][i_s:i_e]).all() and (y[1] == cache["phn"][str(s)][i_s:i_e]).all()
    cache.close()


def test_oracle_spec_window_sampler_split_padding_and_crops(tmp_path):
    """sound_ds.py:262-350: seed-0 train / validation split, zero padding of short utterances, random crops."""
    lens = [50, 12, 33, 8, 70, 41, 10, 29, 64, 55]
    cache = _toy_cache(tmp_path, lens)
    ids = list(range(10))
    trn = list(oracle.spec_window_sampler(cache, ids, 20, batch_size=7, randomize_samples=False, sample_trn=True,
                                          prop_val=0.3, random_seed=5, yield_idxs=True))
    val = list(oracle.spec_window_sampler(cache, ids, 20, batch_size=3, randomize_samples=False, sample_trn=False,
                                          prop_val=0.3, random_seed=5, yield_idxs=True))
    np.random.seed(0)
    perm = np.arange(10)
    np.random.shuffle(perm)
    assert [r[2] for r in trn[0][3]] == list(perm[:-3]) and [r[2] for r in val[0][3]] == list(perm[-3:])
    mfcc, mel, pdb, idx = trn[0]
    assert mfcc.shape == (7, 20, 6) and mel.shape == (7, 20, 5) and pdb.shape == (7, 20, 7)
    for row, (i_s, i_e, s) in enumerate(idx):
        src = cache["power_dB"][str(s)]
        if lens[s] <= 20:
            assert (i_s, i_e) == (0, 20)
            assert (pdb[row, :lens[s]] == src[:]).all() and (pdb[row, lens[s]:] == 0).all()
        else:
            assert (pdb[row] == src[i_s:i_e]).all()
    cache.close()
