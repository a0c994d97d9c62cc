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


# ------------------------------------------------------------------------------------- window samplers on the device
def _wide_cache(tmp_path, lens, widths=(80, 80, 201)):
    """A cache with the hp widths (201 floats per row: windows start on rows that are not 16-byte aligned)."""
    from speech_cloner_b200 import dataset_cache as dc
    path = str(tmp_path / "wide_cache.h5py")
    rng = np.random.default_rng(3)
    w, _ = dc._open_writer(path, "npz")
    with w as out:
        g = {k: out.create_group(k) for k in ("mfcc", "mel_dB", "power_dB", "phn")}
        for i, n in enumerate(lens):
            for k, wd in zip(("mfcc", "mel_dB", "power_dB"), widths):
                g[k].create_dataset(str(i), data=rng.random((n, wd)).astype(np.float32))
            g["phn"].create_dataset(str(i), data=rng.integers(0, 61, size=n).astype(np.int32))
    return dc.open_cache(path)


@pytest.mark.parametrize("widths,n_t", [((80, 80, 201), 40), ((6, 5, 7), 20), ((80, 80, 201), 33), ((3, 1, 2), 7)])
def test_spec_window_sampler_equals_the_reference_loop(al, tmp_path, widths, n_t):
    """Device sampler == literal sound_ds.py:262-350 loop: same [i_s, i_e, i_sample] rows from the same seed, windows
    and zero padding bit-exact, train and validation split, two epochs."""
    from speech_cloner_b200 import dataset_cache as dc
    lens = [50, 12, 133, 8, 70, 41, 10, 29, 64, 155, n_t, n_t + 1, 90, 77]
    cache = _wide_cache(tmp_path, lens, widths)
    dev = dc.DeviceSpecCache.from_cache(cache)
    ids = [i for i in range(len(lens)) if i != 6]
    for sample_trn in (True, False):
        kw = dict(batch_size=4, n_epochs=2, randomize_samples=True, sample_trn=sample_trn, prop_val=0.3, random_seed=17,
                  yield_idxs=True)
        want = list(oracle.spec_window_sampler(cache, ids, n_t, **kw))
        state_after = np.random.get_state()[1].copy()
        got = list(dc.spec_window_sampler(dev, ids, n_t, verbose=False, **kw))
        assert (np.random.get_state()[1] == state_after).all()             # consumed exactly the same random numbers
        assert len(got) == len(want) and len(got) > 0
        padded = 0
        for g, w in zip(got, want):
            assert (g[3] == w[3]).all()
            for a, b in zip(g[:3], w[:3]):
                assert a.is_cuda and a.dtype.is_floating_point and tuple(a.shape) == b.shape
                assert (a.cpu().numpy() == b.astype(np.float32)).all()
            padded += int(sum(lens[s] <= n_t for s in w[3][:, 2]))
        if sample_trn:
            assert padded > 0
    cache.close()


def test_window_sampler_equals_the_reference_loop(al, tmp_path):
    """Device sampler == literal TIMIT_reader.py:474-523 loop (mfcc + phn windows, short utterances skipped)."""
    from speech_cloner_b200 import dataset_cache as dc
    lens = [50, 12, 133, 8, 70, 41, 10, 29, 64, 155, 40, 41, 90, 77]
    cache = _wide_cache(tmp_path, lens)
    dev = dc.DeviceSpecCache.from_cache(cache)
    ids = list(range(len(lens)))
    np.random.seed(9)
    want = list(oracle.window_sampler(cache, ids, 40, batch_size=5, n_epochs=3, yield_idxs=True))
    np.random.seed(9)
    got = list(dc.window_sampler(dev, ids, 40, batch_size=5, n_epochs=3, yield_idxs=True))
    assert len(got) == len(want) and len(got) > 0
    for (x, y, idx), (wx, wy, widx) in zip(got, want):
        assert (idx == widx).all()
        assert tuple(x.shape) == wx.shape and tuple(y.shape) == wy.shape
        assert (x.cpu().numpy() == wx).all() and (y.cpu().numpy() == wy).all()
    cache.close()


def test_device_cache_straight_from_the_front_end(al):
    """A cache that never leaves the GPU: packed front-end outputs wrapped without a copy, windows equal the slices of
    the per-utterance arrays the public API returns; out-of-range windows are cut, not read."""
    import torch
    from speech_cloner_b200 import dataset_cache as dc
    wavs = synth.batch(5, 21, 1.0) + [synth.utterance(2100, 0.2)]
    hp = dict(synth.HP_ENC)
    feats = al.calc_MFCC_input_batch(wavs, **hp)
    layout = al.FrontendLayout([len(w) for w in wavs], hp["hop_length"])
    wav_dev = torch.zeros(layout.total_samples, dtype=torch.float32, device="cuda")
    for w, o in zip(wavs, layout.sample_offsets):
        wav_dev[o:o + len(w)] = torch.from_numpy(w).cuda()
    plan = al._plan_from_kwargs(**hp)
    mfcc, mel, pdb = al.frontend_device(plan, wav_dev, layout)
    dev = dc.DeviceSpecCache.from_device(mfcc, mel, pdb, layout)
    assert dev.groups["power_dB"].data_ptr() == pdb.data_ptr()
    np.random.seed(2)
    batches = list(dc.spec_window_sampler(dev, range(len(wavs)), 60, batch_size=3, prop_val=0.0, yield_idxs=True,
                                          verbose=False))
    assert len(batches) == len(wavs) // 3
    for m, l, p, idx in batches:
        for row, (i_s, i_e, s) in enumerate(idx):
            for got, src in ((m, feats[s][0]), (l, feats[s][1]), (p, feats[s][2])):
                want = np.zeros((60, src.shape[1]), np.float32)
                want[:min(60, src.shape[0] - i_s)] = src[i_s:i_e]
                assert (got[row].cpu().numpy() == want).all()
    # windows outside the packed rows read nothing
    out, = dev.gather(("mel_dB",), [dev.n_rows - 2, dev.n_rows + 5, -1], [60, 60, 60], 60)
    out = out.cpu().numpy()
    assert (out[0, :2] == mel[-2:].cpu().numpy()).all() and (out[0, 2:] == 0).all() and (out[1:] == 0).all()
    # caller-owned output tensors are filled in place; a wrong shape is rejected
    buf = [torch.full((2, 30, 80), 7.0, device="cuda"), torch.full((2, 30, 201), 7.0, device="cuda")]
    res = dev.gather(("mfcc", "power_dB"), [3, 11], [30, 4], 30, out=buf)
    assert res[0].data_ptr() == buf[0].data_ptr() and torch.equal(buf[1][0], pdb[3:33]) and torch.equal(buf[0][1, :4], mfcc[11:15])
    assert float(buf[0][1, 4:].abs().max()) == 0.0
    with pytest.raises(ValueError):
        dev.gather(("mfcc",), [0], [30], 30, out=[torch.empty((1, 31, 80), device="cuda")])



# -------------------------------------------------- against the reference's own caller code (reference_caller_vectors.npz)
def _ref_batches(G, prefix, names):
    out, b = [], 0
    while f"{prefix}/{b}/idxs" in G.files:
        out.append(tuple(G[f"{prefix}/{b}/{n}"] for n in names))
        b += 1
    return out


def test_device_samplers_equal_the_reference_methods(al):
    """Sound_DS.spec_window_sampler / TIMIT.window_sampler executed from the reference files (ast-cut, unmodified) vs
    the device samplers on the same cache, seeds and arguments: same rows, same windows, same random numbers consumed."""
    from speech_cloner_b200 import dataset_cache as dc
    from tests.golden import make_reference_caller_vectors as mc
    G = np.load(mc.OUT)
    cache = mc.sampler_cache("mem://gpu")
    keys = [str(i) for i in range(len(mc.SAMPLER_LENS))]
    dev = dc.DeviceSpecCache.from_arrays({g: [cache[g][k] for k in keys] for g in ("mfcc", "mel_dB", "power_dB", "phn")}, keys)
    ids = np.arange(len(mc.SAMPLER_LENS))[mc.sampler_filter()]
    for r, kw in enumerate(mc.SPEC_RUNS):
        want = _ref_batches(G, f"spec{r}", ("mfcc", "mel_dB", "power_dB", "idxs"))
        got = list(dc.spec_window_sampler(dev, ids, mc.N_TIMESTEPS, random_seed=mc.RANDOM_SEED, yield_idxs=True,
                                          verbose=False, **kw))
        assert len(got) == len(want) > 0
        for g, w in zip(got, want):
            assert np.array_equal(g[3], w[3])
            for a, b in zip(g[:3], w[:3]):
                assert np.array_equal(a.cpu().numpy(), b.astype(np.float32))       # padded reference windows are float64
        assert np.array_equal(np.random.get_state()[1][:8], G[f"spec{r}/rng_after"])
    for r, kw in enumerate(mc.WIN_RUNS):
        want = _ref_batches(G, f"win{r}", ("x", "y", "idxs"))
        np.random.seed(100 + r)
        got = list(dc.window_sampler(dev, ids, mc.N_TIMESTEPS, yield_idxs=True, **kw))
        assert len(got) == len(want) > 0
        for (x, y, idx), (wx, wy, widx) in zip(got, want):
            assert np.array_equal(idx, widx) and np.array_equal(x.cpu().numpy(), wx) and np.array_equal(y.cpu().numpy(), wy)


def test_cache_builder_equals_the_reference_method(al, tmp_path):
    """TIMIT.create_phn_mfcc_cache executed from the reference file (around the reference's audio_lib under the librosa
    shim) vs build_spec_cache on the GPU: same groups and keys, features within 1e-4 / 1e-5, labels exact."""
    from speech_cloner_b200 import dataset_cache as dc
    from tests.golden import make_reference_caller_vectors as mc
    G = np.load(mc.OUT)
    wavs, phn_vs = mc.cache_inputs()
    path = str(tmp_path / "cache.h5py")
    dc.build_spec_cache({"wav": wavs, "phn_v": phn_vs}, mc.CFG, path, phn_conv_d=mc.mr.PHN_CONV, fmt="npz")
    cache = dc.open_cache(path)
    for i in range(len(wavs)):
        for g in ("mfcc", "mel_dB", "power_dB"):
            assert_close(cache[g][str(i)], G[f"cache/{g}/{i}"], what=f"{g}/{i}")
        assert np.array_equal(cache["phn"][str(i)], G[f"cache/phn/{i}"])
    cache.close()


def test_compound_on_device_equals_the_reference_function(al):
    import torch
    from speech_cloner_b200 import conversion as cv
    from tests.golden import make_reference_caller_vectors as mc
    G = np.load(mc.OUT)
    for i, (y0, y1) in enumerate(mc.compound_inputs()):
        got = cv.compound(torch.from_numpy(y0).cuda(), torch.from_numpy(y1).cuda())
        assert np.array_equal(got.cpu().numpy(), G[f"compound{i}/out"])
