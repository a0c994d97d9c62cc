"""Batched dataset-cache builder: the readers' ``for i_sample`` loop on the GPU.

The reference builds its feature caches one utterance at a time:

* ``TIMIT.create_phn_mfcc_cache``   TIMIT_reader.py:144-210
* ``ARCTIC.create_spec_cache``      ARCTIC_reader.py:109-175
* ``TARGET_spk.create_spec_cache``  TARGET_spk_reader.py:132-182 (no phoneme labels)

each iteration calling ``calc_MFCC_input`` (+ ``calc_PHN_target``) and storing four datasets per sample in an HDF5
file with groups ``mfcc/ mel_dB/ power_dB/ phn/`` keyed by ``str(i_sample)``; the file name carries an md5 of the
DSP keys of ``cfg_d`` (TIMIT_reader.py:92-111, ARCTIC_reader.py:62-79, TARGET_spk_reader.py:50-66).

Here the same loop runs in batches through ``sc_frontend_batch`` / ``sc_phn_target_batch`` and writes the same
layout.  ``h5py`` is used when importable (files are then readable by the unchanged reference samplers,
sound_ds.py:262-350); otherwise an ``.npz`` with keys ``"<group>/<i_sample>"`` holds the same arrays
(:func:`open_cache` reads both).  Nothing here computes features on the CPU.
"""
from __future__ import annotations

import hashlib
import os
import sys
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import audio_lib as al

# DSP keys hashed into the cache name, in the reference's order
_KEYS_SPEC = ("sample_rate", "pre_emphasis", "hop_length", "win_length", "n_mels", "n_mfcc", "n_fft", "window",
              "mfcc_normaleze_first_mfcc", "mfcc_norm_factor", "calc_mfcc_derivate", "M_dB_norm_factor",
              "P_dB_norm_factor", "mean_abs_amp_norm", "clip_output")
_KEYS_TIMIT = ("use_all_phonemes",) + _KEYS_SPEC


def spec_cache_name(cfg_d: dict, reader: str = "ARCTIC") -> str:
    """File name the reference derives for a feature cache.

    ``reader="TIMIT"``: md5 over ``use_all_phonemes`` + the DSP keys, base name ``cfg_d['phn_mfcc_cache_name']``
    (TIMIT_reader.py:92-111).  ``"ARCTIC"`` / ``"TARGET"``: DSP keys only, base name ``cfg_d['spec_cache_name']``
    (ARCTIC_reader.py:62-79, TARGET_spk_reader.py:50-66).
    """
    timit = reader.upper() == "TIMIT"
    keys = _KEYS_TIMIT if timit else _KEYS_SPEC
    name_id = hashlib.md5("_".join([str(cfg_d[k]) for k in keys]).encode()).hexdigest()
    base = cfg_d["phn_mfcc_cache_name" if timit else "spec_cache_name"]
    parts = base.split(".")
    return ".".join(parts[:-1]) + "_" + name_id + "." + parts[-1]


def _frontend_kwargs(cfg_d: dict) -> dict:
    """cfg_d -> calc_MFCC_input keyword arguments, the mapping every reader spells out (TIMIT_reader.py:175-190)."""
    return dict(sr=cfg_d["sample_rate"], pre_emphasis=cfg_d["pre_emphasis"], hop_length=cfg_d["hop_length"],
                win_length=cfg_d["win_length"], n_mels=cfg_d["n_mels"], n_mfcc=cfg_d["n_mfcc"], n_fft=cfg_d["n_fft"],
                window=cfg_d["window"], mfcc_normaleze_first_mfcc=cfg_d["mfcc_normaleze_first_mfcc"],
                mfcc_norm_factor=cfg_d["mfcc_norm_factor"], calc_mfcc_derivate=cfg_d["calc_mfcc_derivate"],
                M_dB_norm_factor=cfg_d["M_dB_norm_factor"], P_dB_norm_factor=cfg_d["P_dB_norm_factor"],
                mean_abs_amp_norm=cfg_d["mean_abs_amp_norm"], clip_output=cfg_d["clip_output"])


class _NpzWriter:
    """``create_group(g).create_dataset(str(i), data=a)`` facade over one ``.npz`` (keys ``"g/i"``)."""

    class _Group:
        def __init__(self, store, name):
            self._store, self._name = store, name

        def create_dataset(self, key, data):
            self._store[f"{self._name}/{key}"] = np.asarray(data)

    def __init__(self, path):
        self.path, self._store = path, {}

    def create_group(self, name):
        return self._Group(self._store, name)

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *_):
        if exc_type is None:
            with open(self.path, "wb") as f:       # keep the exact file name (np.savez would append .npz)
                np.savez(f, **self._store)
        return False


def _open_writer(path: str, fmt: str):
    if fmt == "auto":
        try:
            import h5py  # noqa: F401
            fmt = "h5"
        except ImportError:
            fmt = "npz"
    if fmt == "h5":
        import h5py
        return h5py.File(path, "w"), "h5"
    if fmt == "npz":
        return _NpzWriter(path), "npz"
    raise ValueError("fmt must be 'auto', 'h5' or 'npz'")


class CacheReader:
    """Read access with the reference's indexing: ``cache['mfcc'][str(i)][...]`` (sound_ds.py:298-305)."""

    class _Group:
        def __init__(self, npz, name):
            self._npz, self._name = npz, name

        def __getitem__(self, key):
            return self._npz[f"{self._name}/{key}"]

        def __len__(self):
            return sum(1 for k in self._npz.files if k.startswith(self._name + "/"))

        def keys(self):
            return [k.split("/", 1)[1] for k in self._npz.files if k.startswith(self._name + "/")]

    def __init__(self, path):
        self._npz = np.load(path)

    def __getitem__(self, group):
        return self._Group(self._npz, group)

    def close(self):
        self._npz.close()


def open_cache(path: str):
    """Open a cache written by :func:`build_spec_cache` (HDF5 through h5py, or the npz layout)."""
    with open(path, "rb") as f:
        magic = f.read(8)
    if magic.startswith(b"\x89HDF"):
        import h5py
        return h5py.File(path, "r")
    return CacheReader(path)


def _batches(lengths: Sequence[int], max_samples: int, max_utts: int) -> Iterable[range]:
    start, acc = 0, 0
    for i, n in enumerate(lengths):
        if i > start and (acc + n > max_samples or i - start >= max_utts):
            yield range(start, i)
            start, acc = i, 0
        acc += n
    if start < len(lengths):
        yield range(start, len(lengths))


def build_spec_cache(ds: Dict[str, list], cfg_d: dict, path: str, phn_conv_d: Optional[dict] = None,
                     fmt: str = "auto", max_batch_samples: int = 64 * 1024 * 1024, max_batch_utts: int = 4096,
                     verbose: bool = False, overwrite: bool = False) -> Optional[str]:
    """GPU version of ``create_phn_mfcc_cache`` / ``create_spec_cache``.

    ``ds['wav'][i]`` are the float32 waveforms, ``ds['phn_v'][i]`` the ``(start, end, label)`` lists (optional:
    TARGET_spk has none).  Utterances are featurised in batches of at most ``max_batch_samples`` samples
    (256 MB of audio, ~1.4 GB of features on the device by default) and written in sample order under the
    reference's group / key names.  Returns the format written, or None when the file already exists and
    ``overwrite`` is false (the reference warns and returns, TIMIT_reader.py:150-152).
    """
    if os.path.exists(path) and not overwrite:
        print(f' WARNING, build_spec_cache: "{path}" already exists, delete it to rebuild.', file=sys.stderr)
        return None
    wavs: List[np.ndarray] = [np.asarray(w) for w in ds["wav"]]
    phn_vs = ds.get("phn_v") if phn_conv_d is not None else None
    kw = _frontend_kwargs(cfg_d)
    writer, used = _open_writer(path, fmt)
    with writer as out:
        grp = {g: out.create_group(g) for g in ("mfcc", "mel_dB", "power_dB")}
        grp_phn = out.create_group("phn") if phn_vs is not None else None
        for rng in _batches([len(w) for w in wavs], max_batch_samples, max_batch_utts):
            if verbose:
                print(f" - Saved: {rng.start} of {len(wavs)} samples")
            feats = al.calc_MFCC_input_batch([wavs[i] for i in rng], **kw)
            phns = None
            if phn_vs is not None:
                phns = al.calc_PHN_target_batch([len(wavs[i]) for i in rng], [phn_vs[i] for i in rng], phn_conv_d,
                                                hop_length=cfg_d["hop_length"], win_length=cfg_d["win_length"])
            for j, i in enumerate(rng):
                mfcc, mel_db, power_db = feats[j]
                if phns is not None:
                    assert mfcc.shape[0] == phns[j].shape[0], f"sample {i}: mfcc.shape[0] != phn.shape[0]"
                    grp_phn.create_dataset(str(i), data=phns[j])
                grp["mfcc"].create_dataset(str(i), data=mfcc)
                grp["mel_dB"].create_dataset(str(i), data=mel_db)
                grp["power_dB"].create_dataset(str(i), data=power_db)
    return used


# ----------------------------------------------------------------------------------------------- window samplers
# the training-side readers of the cache.  Host-side index logic only (random crops of
# n_timesteps frames); they draw from the global NumPy RNG in exactly the reference's order, so with the same seed and
# the same cache they yield the same batches as TIMIT_reader.py:474-523 / sound_ds.py:262-350.

def window_sampler(cache, sample_ids, n_timesteps, batch_size=32, n_epochs=1, randomize_samples=True, yield_idxs=False):
    """``TIMIT.window_sampler`` (TIMIT_reader.py:474-523): batches of (mfcc window, phn window).

    ``sample_ids``: the utterance indices that pass the reader's ``ds_filter`` (``np.arange(n)[f_s]``).  Utterances
    with ``spec_len <= n_timesteps`` are skipped without consuming random numbers; every kept one draws one
    ``np.random.randint(0, spec_len - n_timesteps)``; each epoch starts with one ``np.random.shuffle``.
    """
    samples_v = [str(int(i)) for i in sample_ids]                 # a Python list, like the reference
    x_v, y_v, idxs_v = [], [], []
    for _ in range(n_epochs):
        if randomize_samples:
            np.random.shuffle(samples_v)
        for i_sample in samples_v:
            mfcc = cache["mfcc"][i_sample]
            spec_len = mfcc.shape[0]
            if spec_len <= n_timesteps:
                continue
            i_s = np.random.randint(0, spec_len - n_timesteps)
            i_e = i_s + n_timesteps
            x_v.append(mfcc[i_s:i_e])
            y_v.append(cache["phn"][i_sample][i_s:i_e])
            idxs_v.append([i_s, i_e, int(i_sample)])
            if len(x_v) == batch_size:
                x, y = np.array(x_v), np.array(y_v)
                assert x.shape[1] == y.shape[1] == n_timesteps
                yield (x, y, np.array(idxs_v)) if yield_idxs else (x, y)
                x_v, y_v, idxs_v = [], [], []


def spec_window_sampler(cache, sample_ids, n_timesteps, batch_size=32, n_epochs=1, randomize_samples=True,
                        sample_trn=True, prop_val=0.3, random_seed=None, yield_idxs=False, verbose=True):
    """``Sound_DS.spec_window_sampler`` (sound_ds.py:262-350): batches of (mfcc, mel_dB, power_dB) windows.

    The train / validation split is the reference's: seed 0, shuffle of ``arange``, last ``int(prop_val * n)`` ids are
    validation, then ``np.random.seed(random_seed)``.  Utterances with ``spec_len <= n_timesteps`` are zero padded (as
    float64, like ``_zero_pad``'s ``np.zeros``) instead of cropped and consume no random number.
    """
    samples_v = np.array([str(int(i)) for i in sample_ids])       # a NumPy array of str, like the reference
    if prop_val > 0.0:
        np.random.seed(0)
        idx_v = np.arange(samples_v.shape[0])
        np.random.shuffle(idx_v)
        n_val = int(prop_val * samples_v.shape[0])
        samples_v = samples_v[idx_v[:-n_val]] if sample_trn else samples_v[idx_v[-n_val:]]
        np.random.seed(random_seed)
    names = ("mfcc", "mel_dB", "power_dB")
    acc = {k: [] for k in names}
    idxs_v, n_warning = [], 0
    for _ in range(n_epochs):
        if randomize_samples:
            np.random.shuffle(samples_v)
        for i_sample in samples_v:
            spec_len = cache["mfcc"][i_sample].shape[0]
            if spec_len <= n_timesteps:
                i_s, i_e = 0, n_timesteps
                pad = n_timesteps - spec_len
                for k in names:
                    a = cache[k][i_sample][:]
                    acc[k].append(np.concatenate([a, np.zeros((pad, a.shape[1]))], axis=0))
                if verbose and n_warning < 5:
                    print("WARNING: padding!!!")
                    n_warning += 1
            else:
                i_s = np.random.randint(0, spec_len - n_timesteps)
                i_e = i_s + n_timesteps
                for k in names:
                    acc[k].append(cache[k][i_sample][i_s:i_e])
            idxs_v.append([i_s, i_e, int(i_sample)])
            if len(acc["mfcc"]) == batch_size:
                out = tuple(np.array(acc[k]) for k in names)
                assert out[0].shape[1] == out[1].shape[1] == out[2].shape[1] == n_timesteps
                yield out + (np.array(idxs_v),) if yield_idxs else out
                acc = {k: [] for k in names}
                idxs_v = []
