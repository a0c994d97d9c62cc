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
