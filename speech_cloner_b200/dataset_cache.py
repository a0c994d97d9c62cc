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

The training-side readers of the cache are here too: :class:`DeviceSpecCache` keeps a cache
resident in device memory and :func:`spec_window_sampler` / :func:`window_sampler` cut the reference's random windows
out of it with one ``sc_window_gather`` launch per batch (sound_ds.py:262-350, TIMIT_reader.py:474-523).
"""
from __future__ import annotations

import hashlib
import os
import sys
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import _lib
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
# SURVEY.md §8(f) rank 4.  The reference's samplers re-open the HDF5 cache and slice one window per utterance on the host
# (`ds_h5py['mfcc'][i_sample][i_s:i_e]`, one h5py read per group and window).  Here the cache stays in HBM (10 h of audio
# are 10.4 GB of features) and a batch is ONE kernel launch; only the random numbers stay on the host, drawn from NumPy's
# global generator in exactly the reference's order, so the same seed gives the same windows.

class DeviceSpecCache:
    """A feature cache resident in device memory: packed ``[rows][width]`` CUDA tensors per group.

    ``groups``        ``{"mfcc": …, "mel_dB": …, "power_dB": …[, "phn": …]}`` (float32; ``phn`` int32, ``[rows]`` or
                      ``[rows][k]`` for one-hot targets), all sharing one row layout
    ``frame_offsets`` first row of every utterance, ``spec_len`` its number of frames
    ``keys``          the cache key (``str(i_sample)``) of every slot
    """

    FEATURES = ("mfcc", "mel_dB", "power_dB")

    def __init__(self, groups, frame_offsets, spec_len, keys=None):
        torch = al._require_cuda()
        self.groups = {}
        rows = None
        for name, t in groups.items():
            if not (al._is_tensor(t) and t.is_cuda and t.is_contiguous() and t.element_size() == 4):
                raise ValueError(f"{name}: contiguous CUDA tensor of 32-bit elements expected")
            if rows is None:
                rows = t.shape[0]
            elif t.shape[0] != rows:
                raise ValueError("all groups must share the row layout")
            self.groups[name] = t
        self.n_rows = int(rows)
        self.frame_offsets = np.asarray(frame_offsets, dtype=np.int64)
        self.spec_len = np.asarray(spec_len, dtype=np.int64)
        if self.frame_offsets.shape != self.spec_len.shape:
            raise ValueError("frame_offsets and spec_len differ in length")
        if len(self.spec_len) and int((self.frame_offsets + self.spec_len).max()) > self.n_rows:
            raise ValueError("an utterance reaches beyond the packed buffers")
        keys = [str(i) for i in range(len(self.spec_len))] if keys is None else [str(k) for k in keys]
        self.slot = {k: i for i, k in enumerate(keys)}
        self._torch = torch

    # -- constructors
    @classmethod
    def from_device(cls, mfcc, mel_dB, power_dB, layout, phn=None, keys=None):
        """Wrap the packed outputs of :func:`audio_lib.frontend_device` (no copy): a cache that never left the GPU."""
        groups = {"mfcc": mfcc, "mel_dB": mel_dB, "power_dB": power_dB}
        if phn is not None:
            groups["phn"] = phn
        return cls(groups, layout.frame_offsets[:-1], layout.frames, keys)

    @classmethod
    def from_arrays(cls, arrays: Dict[str, Sequence[np.ndarray]], keys=None):
        """Upload per-utterance host arrays (``arrays[group][u]`` of shape ``[T_u, width]`` or ``[T_u]``)."""
        torch = al._require_cuda()
        names = list(arrays)
        lens = [int(np.shape(a)[0]) for a in arrays[names[0]]]
        off = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        groups = {}
        for name in names:
            seq = arrays[name]
            if [int(np.shape(a)[0]) for a in seq] != lens:
                raise ValueError(f"{name}: frame counts differ from {names[0]}")
            first = np.asarray(seq[0])
            dt = np.int32 if np.issubdtype(first.dtype, np.integer) else np.float32
            host = torch.empty((int(off[-1]),) + first.shape[1:], dtype=torch.int32 if dt is np.int32 else torch.float32,
                               pin_memory=True)
            view = host.numpy()
            for u, a in enumerate(seq):
                view[off[u]:off[u + 1]] = np.asarray(a)
            groups[name] = host.to("cuda", non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return cls(groups, off[:-1], lens, keys)

    @classmethod
    def from_cache(cls, cache, sample_ids=None, groups=None):
        """Load an open cache (:func:`open_cache` / h5py): all utterances, or the ones in ``sample_ids``."""
        if sample_ids is None:
            sample_ids = sorted((int(k) for k in cache["mfcc"].keys()))
        keys = [str(int(i)) for i in sample_ids]
        if groups is None:
            groups = [g for g in cls.FEATURES + ("phn",) if _has_group(cache, g)]
        return cls.from_arrays({g: [cache[g][k][...] for k in keys] for g in groups}, keys)

    # -- one batch
    def gather(self, names, first_rows, valid_rows, n_timesteps, out=None):
        """Windows ``[first_rows[w], +valid_rows[w])`` of the packed rows, zero padded to ``n_timesteps`` rows, from every
        group in ``names``: dense ``[n_windows, n_timesteps(, width)]`` CUDA tensors out of one launch (written into
        ``out``, a matching list of tensors, when given)."""
        torch = self._torch
        lib = _lib.load()
        n_w = len(first_rows)
        stage = torch.empty(3 * n_w, dtype=torch.int32, pin_memory=True)     # int64 first rows + int32 valid counts
        host = stage.numpy()
        host[:2 * n_w].view(np.int64)[:] = np.asarray(first_rows, dtype=np.int64)
        host[2 * n_w:] = np.asarray(valid_rows, dtype=np.int32)
        dev = stage.to("cuda", non_blocking=True)
        outs, srcs, dsts, widths = [], [], [], []
        for a, name in enumerate(names):
            t = self.groups[name]
            shape = (n_w, int(n_timesteps)) + tuple(t.shape[1:])
            if out is None:
                o = torch.empty(shape, dtype=t.dtype, device="cuda")
            else:
                o = out[a]
                if tuple(o.shape) != shape or o.dtype != t.dtype or not (o.is_cuda and o.is_contiguous()):
                    raise ValueError(f"out[{a}]: contiguous CUDA tensor of shape {shape} and dtype {t.dtype} expected")
            outs.append(o)
            srcs.append(t.data_ptr()); dsts.append(o.data_ptr()); widths.append(int(np.prod(t.shape[1:], dtype=np.int64)))
        for a0 in range(0, len(names), 4):
            k = len(names[a0:a0 + 4])
            rc = lib.sc_window_gather((_lib.C.c_void_p * k)(*srcs[a0:a0 + k]), (_lib.C.c_void_p * k)(*dsts[a0:a0 + k]),
                                      _lib.i64_array(widths[a0:a0 + k]), k, self.n_rows, dev.data_ptr(),
                                      dev.data_ptr() + 8 * n_w, n_w, int(n_timesteps), al._stream_ptr(torch))
            _lib.check(rc, "sc_window_gather")
        return outs


def _has_group(cache, name) -> bool:
    try:
        cache[name]
        return len(cache[name]) > 0
    except KeyError:
        return False


class _BatchPlan:
    """Rows of one batch under construction: first packed row, valid rows and the reference's index triple per window."""

    def __init__(self, cache: "DeviceSpecCache", n_timesteps: int, batch_size: int):
        self.cache, self.n_t, self.size = cache, int(n_timesteps), int(batch_size)
        self.clear()

    def clear(self):
        self.first_rows, self.valid_rows, self.triples = [], [], []

    def add(self, slot: int, start: int, valid: int, sample_id: int) -> bool:
        """Append the window ``[start, start + n_timesteps)`` of utterance ``slot``; True when the batch is full."""
        self.first_rows.append(int(self.cache.frame_offsets[slot]) + start)
        self.valid_rows.append(valid)
        self.triples.append([start, start + self.n_t, sample_id])
        return len(self.first_rows) == self.size

    def launch(self, names, yield_idxs: bool):
        tensors = tuple(self.cache.gather(names, self.first_rows, self.valid_rows, self.n_t))
        out = tensors + (np.array(self.triples),) if yield_idxs else tensors
        self.clear()
        return out


def _reference_split(keys: np.ndarray, prop_val: float, sample_trn: bool, random_seed) -> np.ndarray:
    """The train / validation split of sound_ds.py:269-284: a seed-0 permutation whose last ``int(prop_val * n)`` entries
    are the validation set, then the generator is re-seeded with the reader's ``random_seed``."""
    if prop_val <= 0.0:
        return keys
    np.random.seed(0)
    perm = np.arange(keys.shape[0])
    np.random.shuffle(perm)
    n_val = int(prop_val * keys.shape[0])
    picked = keys[perm[:-n_val]] if sample_trn else keys[perm[-n_val:]]
    np.random.seed(random_seed)
    return picked


def spec_window_sampler(cache: DeviceSpecCache, sample_ids, n_timesteps, batch_size=32, n_epochs=1,
                        randomize_samples=True, sample_trn=True, prop_val=0.3, random_seed=None, yield_idxs=False,
                        verbose=True):
    """``Sound_DS.spec_window_sampler`` (sound_ds.py:262-350) on a device-resident cache.

    Yields ``(mfcc_v, mel_dB_v, power_dB_v[, idxs_v])``: float32 CUDA tensors ``[batch_size, n_timesteps, width]`` and
    the reference's ``[i_s, i_e, i_sample]`` rows as a NumPy array.  ``sample_ids`` is what the reader's filter selects
    (``np.arange(n)[f_s]``), ``random_seed`` the reader's ``self.random_seed``.  Same random stream as the reference: the
    seed-0 train / validation split, ``np.random.seed(random_seed)``, one ``np.random.shuffle`` per epoch, one
    ``np.random.randint(0, spec_len - n_timesteps)`` per utterance longer than the window; shorter ones are zero padded
    without a draw.  (The reference's padded windows are float64 because of ``np.zeros``; the values are the same.)
    """
    # a NumPy array of str keys: np.random.shuffle must permute the same container type as the reference's
    keys = _reference_split(np.array([str(int(i)) for i in sample_ids]), prop_val, sample_trn, random_seed)
    plan = _BatchPlan(cache, n_timesteps, batch_size)
    padded_notes = 0
    for _ in range(n_epochs):
        if randomize_samples:
            np.random.shuffle(keys)
        for key in keys:
            slot = cache.slot[str(key)]
            frames = int(cache.spec_len[slot])
            if frames > n_timesteps:
                start, valid = np.random.randint(0, frames - n_timesteps), n_timesteps
            else:                                                 # zero padded, no random number consumed (:301-311)
                start, valid = 0, frames
                if verbose and padded_notes < 5:
                    print("WARNING: padding!!!")
                    padded_notes += 1
            if plan.add(slot, start, valid, int(key)):
                yield plan.launch(DeviceSpecCache.FEATURES, yield_idxs)


def window_sampler(cache: DeviceSpecCache, sample_ids, n_timesteps, batch_size=32, n_epochs=1, randomize_samples=True,
                   yield_idxs=False):
    """``TIMIT.window_sampler`` (TIMIT_reader.py:474-523) on a device-resident cache: ``(x_v, y_v[, idxs_v])`` with the
    mfcc windows ``[batch_size, n_timesteps, width]`` and the phoneme targets ``[batch_size, n_timesteps(, k)]`` as CUDA
    tensors.  Utterances with ``spec_len <= n_timesteps`` are skipped without consuming a random number."""
    keys = [str(int(i)) for i in sample_ids]                      # a Python list here: the reference shuffles a list (:478)
    plan = _BatchPlan(cache, n_timesteps, batch_size)
    for _ in range(n_epochs):
        if randomize_samples:
            np.random.shuffle(keys)
        for key in keys:
            slot = cache.slot[key]
            frames = int(cache.spec_len[slot])
            if frames <= n_timesteps:
                continue
            if plan.add(slot, np.random.randint(0, frames - n_timesteps), n_timesteps, int(key)):
                yield plan.launch(("mfcc", "phn"), yield_idxs)
