"""Run the reference's own CALLER code either side of the hot path and freeze what it produces in
reference_caller_vectors.npz:

    python tests/golden/make_reference_caller_vectors.py

* ``compound``                          test.py:46-84          window stitching after decoder.predict
* ``TIMIT.create_phn_mfcc_cache``       TIMIT_reader.py:140-210 the per-utterance cache loop (calls calc_MFCC_input +
                                                                calc_PHN_target of the reference's audio_lib)
* ``Sound_DS.spec_window_sampler``      sound_ds.py:246-350     random windows + zero padding, train / validation split
* ``TIMIT.window_sampler``              TIMIT_reader.py:474-523 (mfcc, phn) windows

The modules themselves need TensorFlow / h5py / librosa / a dataset on disk, so each function is cut out of its file with
``ast`` and executed UNMODIFIED in a namespace that provides NumPy, an in-memory stand-in for ``h5py.File`` and the
reference's audio_lib (imported under the librosa shim, see make_reference_vectors.py); methods get a plain object as
``self`` with the few attributes they read.  Nothing is copied into the repository: the source text is read from
/root/reference at generation time.
"""
import ast
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from speech_cloner_b200 import synth  # noqa: E402
from tests.golden import make_reference_vectors as mr  # noqa: E402

REF_DIR = "/root/reference"
OUT = os.path.join(HERE, "reference_caller_vectors.npz")

CFG = dict(use_all_phonemes=True, sample_rate=16000, pre_emphasis=0.97, hop_length=80, win_length=400, n_mels=80,
           n_mfcc=40, n_fft=None, window="hann", mfcc_normaleze_first_mfcc=True, mfcc_norm_factor=0.01,
           calc_mfcc_derivate=True, M_dB_norm_factor=0.01, P_dB_norm_factor=0.01, mean_abs_amp_norm=0.003,
           clip_output=True)
SAMPLER_LENS = [50, 12, 33, 8, 70, 41, 10, 29, 64, 55, 20, 21, 47]
SAMPLER_WIDTHS = {"mfcc": 6, "mel_dB": 5, "power_dB": 7}
N_TIMESTEPS = 20


# ------------------------------------------------------------------------------- in-memory h5py.File stand-in
class MemFile:
    store = {}                                        # path -> {group: {key: array}}

    class _Group(dict):
        def create_dataset(self, key, data):
            self[key] = np.asarray(data)

    def __init__(self, path, mode="r"):
        if mode == "w":
            MemFile.store[path] = {}
        self._groups = MemFile.store[path]

    def create_group(self, name):
        return self._groups.setdefault(name, MemFile._Group())

    def __getitem__(self, name):
        return self._groups[name]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def cut(file_name, func, cls=None):
    """Source text of a top-level function or of a method, exactly as written in the reference file."""
    text = open(os.path.join(REF_DIR, file_name)).read()
    tree = ast.parse(text)
    body = tree.body
    if cls is not None:
        body = next(n for n in body if isinstance(n, ast.ClassDef) and n.name == cls).body
    node = next(n for n in body if isinstance(n, ast.FunctionDef) and n.name == func)
    src = "\n".join(text.splitlines()[node.lineno - 1:node.end_lineno])
    return src if cls is None else "class _Holder:\n" + src          # a method keeps its indentation inside a holder class


def load(file_name, func, cls=None, **extra):
    ns = dict(np=np, os=os, sys=sys, h5py=types.SimpleNamespace(File=MemFile))
    ns.update(extra)
    exec(compile(cut(file_name, func, cls), f"{REF_DIR}/{file_name}:{func}", "exec"), ns)
    return ns[func] if cls is None else ns["_Holder"].__dict__[func]


# ----------------------------------------------------------------------------------------------- seeded inputs
def compound_inputs():
    rng = np.random.default_rng(21)
    return [(rng.random((5, 8, 3)), rng.random((4, 8, 3))), (rng.random((1, 12, 2)), rng.random((0, 12, 2))),
            (rng.random((2, 16, 4)), rng.random((1, 16, 4)))]


def cache_inputs():
    """Three short utterances with phoneme intervals; seeds whose gain does not depend on the NumPy generation."""
    rng = np.random.default_rng(31)
    wavs, phn_vs = [], []
    for seed0, seconds in ((8000, 0.31), (8100, 0.44), (8200, 0.2)):
        _, y = mr.fe_input(seed0, seconds, (0.0, 10.0), CFG["mean_abs_amp_norm"])
        cuts = np.sort(rng.choice(np.arange(1, len(y)), size=4, replace=False))
        b = [0] + [int(c) for c in cuts] + [len(y)]
        wavs.append(y)
        phn_vs.append([(b[i], b[i + 1], "abcd"[i % 4]) for i in range(len(b) - 1)])
    return wavs, phn_vs


def sampler_cache(path="mem://toy"):
    rng = np.random.default_rng(41)
    f = MemFile(path, "w")
    groups = {g: f.create_group(g) for g in list(SAMPLER_WIDTHS) + ["phn"]}
    for i, n in enumerate(SAMPLER_LENS):
        for g, w in SAMPLER_WIDTHS.items():
            groups[g].create_dataset(str(i), rng.random((n, w)).astype(np.float32))
        groups["phn"].create_dataset(str(i), rng.integers(0, 61, size=n).astype(np.int32))
    return f


def sampler_filter():
    f_s = np.ones(len(SAMPLER_LENS), dtype=bool)
    f_s[6] = False
    return f_s


SPEC_RUNS = [dict(batch_size=4, n_epochs=2, randomize_samples=True, sample_trn=True, prop_val=0.3),
             dict(batch_size=3, n_epochs=1, randomize_samples=False, sample_trn=False, prop_val=0.3),
             dict(batch_size=5, n_epochs=2, randomize_samples=True, sample_trn=True, prop_val=0.0)]
WIN_RUNS = [dict(batch_size=3, n_epochs=2, randomize_samples=True), dict(batch_size=4, n_epochs=1, randomize_samples=False)]
RANDOM_SEED = 17


def compute():
    out = {}
    # compound (test.py:46-84)
    compound = load("test.py", "compound")
    for i, (y0, y1) in enumerate(compound_inputs()):
        out[f"compound{i}/out"] = compound(y0, y1)

    # the cache loop (TIMIT_reader.py:140-210) around the reference's audio_lib
    ref = mr.load_reference()
    create = load("TIMIT_reader.py", "create_phn_mfcc_cache", "TIMIT", calc_MFCC_input=ref.calc_MFCC_input,
                  calc_PHN_target=ref.calc_PHN_target)
    wavs, phn_vs = cache_inputs()
    me = types.SimpleNamespace(cfg_d=CFG, ds_path="mem://", spec_cache_name="cache", phn2ohv=mr.PHN_CONV,
                               ds={"wav": wavs, "phn_v": phn_vs}, verbose=False)
    stdout, sys.stdout = sys.stdout, open(os.devnull, "w")
    try:
        create(me)
    finally:
        sys.stdout = stdout
    written = MemFile.store[os.path.join("mem://", "cache")]
    for g in ("mfcc", "mel_dB", "power_dB", "phn"):
        for k, v in written[g].items():
            out[f"cache/{g}/{k}"] = v

    # samplers (sound_ds.py:246-350, TIMIT_reader.py:474-523)
    sampler_cache(os.path.join("mem://", "toy"))
    zero_pad = load("sound_ds.py", "_zero_pad", "Sound_DS")
    me = types.SimpleNamespace(n_timesteps=N_TIMESTEPS, ds_path="mem://", spec_cache_name="toy", random_seed=RANDOM_SEED,
                               get_ds_filter=lambda d: sampler_filter())
    me._zero_pad = lambda *a, **k: zero_pad(me, *a, **k)
    spec = load("sound_ds.py", "spec_window_sampler", "Sound_DS")
    stdout, sys.stdout = sys.stdout, open(os.devnull, "w")
    try:
        for r, kw in enumerate(SPEC_RUNS):
            for b, (mfcc, mel, pdb, idxs) in enumerate(spec(me, yield_idxs=True, **kw)):
                out[f"spec{r}/{b}/mfcc"], out[f"spec{r}/{b}/mel_dB"], out[f"spec{r}/{b}/power_dB"] = mfcc, mel, pdb
                out[f"spec{r}/{b}/idxs"] = idxs
            out[f"spec{r}/rng_after"] = np.random.get_state()[1][:8].copy()
    finally:
        sys.stdout = stdout
    win = load("TIMIT_reader.py", "window_sampler", "TIMIT")
    for r, kw in enumerate(WIN_RUNS):
        np.random.seed(100 + r)
        for b, (x, y, idxs) in enumerate(win(me, ds_filter_d={}, yield_idxs=True, **kw)):
            out[f"win{r}/{b}/x"], out[f"win{r}/{b}/y"], out[f"win{r}/{b}/idxs"] = x, y, idxs
    return out


def main():
    out = compute()
    np.savez_compressed(OUT, **out)
    print(len(out), "arrays;", sorted({k.split("/")[0] for k in out}))


if __name__ == "__main__":
    main()
