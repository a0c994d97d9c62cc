"""GPU parity of the front-end (calc_MFCC_input, /root/reference/audio_lib.py:89-244) against the oracle.

Tolerance (BASELINE.json north_star): 1e-4 relative / 1e-5 absolute on all three float32 outputs.
Every call goes through the C ABI of include/speechdsp.h via speech_cloner_b200.audio_lib.
"""
import numpy as np
import pytest

from oracle import audio_lib_oracle as oracle
from speech_cloner_b200 import synth
from tests.util import assert_close

pytestmark = pytest.mark.gpu

HP = dict(synth.HP_ENC)


@pytest.fixture(scope="module")
def al(built_lib):
    from speech_cloner_b200 import audio_lib
    return audio_lib


def _check(al, y, what, atol=1e-5, **kw):
    got = al.calc_MFCC_input(y, **kw)
    want = oracle.calc_MFCC_input(y, **kw)
    for g, w, name in zip(got, want, ("MFCC", "M_dB", "P_dB")):
        assert g.dtype == np.float32 and g.flags["C_CONTIGUOUS"]
        assert_close(g, w, atol=atol, what=f"{what}/{name}")


def test_single_utterance_hp(al):
    _check(al, synth.utterance(1000, 3.0, ds_norm=(0.0, 10.0)), "hp 3s", **HP)


def test_config1_batch_32x3s(al):
    """BASELINE.json configs[0]: 32 synthetic 16 kHz 3 s TIMIT-shaped utterances at hp/ds_enc_cfg_d.json."""
    wavs = synth.batch(1, 32, 3.0, ds_norm=(0.0, 10.0))
    got = al.calc_MFCC_input_batch(wavs, **HP)
    assert len(got) == 32
    for i, (y, g) in enumerate(zip(wavs, got)):
        want = oracle.calc_MFCC_input(y, **HP)
        assert g[0].shape == (601, 80) and g[1].shape == (601, 80) and g[2].shape == (601, 201)
        for a, b, name in zip(g, want, ("MFCC", "M_dB", "P_dB")):
            assert_close(a, b, what=f"utt {i}/{name}")


def test_ragged_batch_equals_single_calls(al):
    lens = [48000, 16001, 7999, 400, 399, 81, 80, 160, 33333]
    wavs = [synth.utterance(50 + i, n / 16000.0)[:n] for i, n in enumerate(lens)]
    got = al.calc_MFCC_input_batch(wavs, **HP)
    for y, g in zip(wavs, got):
        single = al.calc_MFCC_input(y, **HP)
        want = oracle.calc_MFCC_input(y, **HP)
        for a, s, w in zip(g, single, want):
            assert a.shape == w.shape == s.shape
            np.testing.assert_array_equal(a, s)          # batching must not change a bit
            assert_close(a, w, what=f"len {len(y)}")


@pytest.mark.parametrize("case", ["noise", "sine_on_bin", "sine_between_bins", "impulse", "leading_zeros", "short"])
def test_adversarial_inputs(al, case):
    rng = np.random.default_rng(7)
    n = 16000
    t = np.arange(n) / 16000.0
    if case == "noise":
        y = 0.1 * rng.standard_normal(n)
    elif case == "sine_on_bin":
        y = 0.1 * np.sin(2 * np.pi * 1000.0 * t) + 1e-4 * rng.standard_normal(n)
    elif case == "sine_between_bins":
        y = 0.1 * np.sin(2 * np.pi * 1020.0 * t) + 1e-4 * rng.standard_normal(n)
    elif case == "impulse":
        y = 1e-4 * rng.standard_normal(n); y[5000] = 0.5
    elif case == "leading_zeros":
        y = 0.05 * rng.standard_normal(n); y[:4000] = 0.0
    else:
        y = 0.1 * rng.standard_normal(250)                   # L < n_fft: reflect pad wraps more than once
    _check(al, y.astype(np.float32), case, **HP)


@pytest.mark.parametrize("override", [
    dict(calc_mfcc_derivate=False),
    dict(mfcc_normaleze_first_mfcc=False),
    dict(clip_output=False),
    dict(pre_emphasis=0.0),
    dict(mean_abs_amp_norm=1.0),
    dict(P_dB_norm_factor=1.0, M_dB_norm_factor=1.0, mfcc_norm_factor=1.0, clip_output=False),
    dict(window="hamming"),
    dict(win_length=320, n_fft=400),
    dict(n_mels=128, n_mfcc=20),
    dict(n_mels=40, n_mfcc=13, sr=8000),
])
def test_parameter_switches(al, override):
    kw = dict(HP); kw.update(override)
    # the 1e-5 absolute tolerance is stated for the hp-normalised outputs (x 0.01); un-normalised dB
    # outputs carry the same error 100x larger, i.e. 1e-3 dB
    atol = 1e-3 if kw["P_dB_norm_factor"] == 1.0 else 1e-5
    _check(al, synth.utterance(77, 1.5), str(override), atol=atol, **kw)


@pytest.mark.parametrize("geom", [
    dict(hop_length=40, win_length=400, n_mels=128),           # calc_MFCC_input's own defaults (:92-94)
    dict(hop_length=128, win_length=512, n_fft=512),
    dict(hop_length=100, win_length=300, n_fft=384, window="hamming"),
])
def test_generic_geometry(al, geom):
    kw = dict(HP); kw.update(geom)
    _check(al, synth.utterance(78, 1.0), str(geom), **kw)


@pytest.mark.parametrize("case", ["speech", "sine"])
def test_fp32_fast_mode_documented_accuracy(al, case):
    """Opt-in float32 FFT: everything within 1e-4 absolute, and >= 99.7 % of the elements within the
    float64 mode's tolerance; the stragglers are bins 70-80 dB below the utterance maximum."""
    if case == "speech":
        y = synth.utterance(1000, 3.0, ds_norm=(0.0, 10.0))
    else:
        t = np.arange(16000) / 16000.0
        y = (0.1 * np.sin(2 * np.pi * 1000.0 * t) + 1e-4 * np.random.default_rng(7).standard_normal(16000)).astype(np.float32)
    got = al.calc_MFCC_input_batch([y], fft_precision="fp32", **HP)[0]
    want = oracle.calc_MFCC_input(y, **HP)
    for g, w in zip(got, want):
        err = np.abs(g.astype(np.float64) - w)
        assert err.max() < 1e-4
        assert np.mean(err <= 1e-5 + 1e-4 * np.abs(w)) >= 0.997


def test_gain_matches_numpy_bitwise(al):
    """np.abs(y).mean() bit for bit, through the SAME kernels the hot path launches (k_fe_setup -> k_abs_pairwise4 ->
    k_gain_finalize): sc_mean_abs_batch and sc_frontend_batch share them since round 2."""
    """mean|y| must be NumPy's float32 pairwise sum bit for bit (a 1-ulp gain moves near-floor bins by 3e-5)."""
    import torch
    rng = np.random.default_rng(3)
    lens = [1, 5, 8, 9, 127, 128, 129, 255, 1000, 4097, 7999, 8000, 8001, 16001, 48000, 64000, 100003, 1 << 20]
    # lengths whose largest sub-tree exceeds 8 000 samples (NumPy's split rounds the left half down to a multiple of 8):
    # 8 001 .. 8 015 samples in one staged sub-tree
    lens += [15993, 15998, 15999, 31977, 31999, 63998, 63999, 127999, 255985, 8000 * 64 - 1]
    wavs = [(0.1 * rng.standard_normal(n)).astype(np.float32) for n in lens]
    plan = al.DspPlan.get(n_fft=400, win_length=400, hop_length=80)
    lay = al.FrontendLayout(lens, 80)
    dev = torch.zeros(lay.total_samples, dtype=torch.float32, device="cuda")
    for w, o in zip(wavs, lay.sample_offsets):
        dev[o:o + len(w)] = torch.from_numpy(w).cuda()
    got = al.mean_abs_device(plan, dev, lay).cpu().numpy()
    want = np.array([np.abs(w).mean() for w in wavs], dtype=np.float32)
    np.testing.assert_array_equal(got, want)


def test_lengths_just_below_a_subtree_multiple(al):
    """Regression (scripts/soak.py): utterances of 2^D * 8000 - 1 .. - 7 samples stage a sub-tree of up to 8 015 samples in the
    |y| kernel; the gain, and with it every feature above the amin clamp, must still match."""
    rng = np.random.default_rng(21)
    wavs = [(0.05 * rng.standard_normal(n)).astype(np.float32) for n in (15999, 31999, 15993)] + [synth.utterance(11, 4.0)[:63999]]
    for y, g in zip(wavs, al.calc_MFCC_input_batch(wavs, **HP)):
        for a, b, name in zip(g, oracle.calc_MFCC_input(y, **HP), ("mfcc", "mel", "pdb")):
            assert_close(a, b, what=f"{name} (n = {len(y)})")


def test_invalid_inputs_raise(al):
    with pytest.raises(ValueError):
        al.calc_MFCC_input([0.0, 1.0], **HP)
    with pytest.raises(ValueError):
        al.calc_MFCC_input(np.zeros(100, dtype=np.int16), **HP)
    with pytest.raises(ValueError):
        al.calc_MFCC_input(np.zeros((2, 100), dtype=np.float32), **HP)
    y = np.ones(1000, dtype=np.float32); y[3] = np.nan
    with pytest.raises(ValueError):
        al.calc_MFCC_input(y, **HP)
    with pytest.raises(ValueError):
        al.calc_MFCC_input(np.ones(40, dtype=np.float32), **HP)      # 1 frame + delta -> reference raises too


def test_device_tensor_in_out(al):
    import torch
    y = synth.utterance(5, 2.0)
    want = oracle.calc_MFCC_input(y, **HP)
    got = al.calc_MFCC_input(torch.from_numpy(y).cuda(), **HP)
    for g, w in zip(got, want):
        assert g.is_cuda
        assert_close(g.cpu().numpy(), w)


def test_preemphasis_pair(al):
    y = synth.utterance(9, 1.0)
    pe = al.calc_preemphasis(y, 0.97)
    assert pe.dtype == np.float64
    np.testing.assert_allclose(pe, oracle.calc_preemphasis(y, 0.97), rtol=0, atol=1e-12)
    inv = al.calc_inv_preemphasis(y, 0.97)
    assert inv.dtype == np.float64
    np.testing.assert_allclose(inv, oracle.calc_inv_preemphasis(y, 0.97), rtol=1e-12, atol=1e-13)
    # round trip through both filters recovers the float32 signal
    np.testing.assert_allclose(al.calc_inv_preemphasis(pe.astype(np.float32), 0.97), y, atol=2e-6)


def test_long_utterance_many_tiles(al):
    """One 60 s utterance: 500 tiles of the warp-specialised pass A and a depth-7 |y| tree in a single utterance."""
    y = np.concatenate([synth.utterance(90 + i, 10.0) for i in range(6)]).astype(np.float32)
    _check(al, y, "60 s", **HP)


def test_tightly_packed_unaligned_offsets(al):
    """The C ABI does not require 16-byte aligned utterance starts: packed offsets with odd lengths take the
    4-byte staging path of pass A and the scalar pass B, and must give the same numbers as the aligned layout."""
    import torch
    lens = [16001, 9999, 4497, 31237]
    wavs = [synth.utterance(60 + i, n / 16000.0)[:n] for i, n in enumerate(lens)]
    kw = dict(sr=16000, n_fft=400, win_length=400, hop_length=80, n_mels=80, n_mfcc=40, window="hann", pre_emphasis=0.97,
              mfcc_normaleze_first_mfcc=True, mfcc_norm_factor=0.01, calc_mfcc_derivate=True, M_dB_norm_factor=0.01,
              P_dB_norm_factor=0.01, mean_abs_amp_norm=0.003, clip_output=True)
    plan = al.DspPlan(**kw)
    lay = al.FrontendLayout(lens, 80)
    so, fo = [0], [0]
    for n, t in zip(lens, lay.frames):
        so.append(so[-1] + n)
        fo.append(fo[-1] + t)
    lay.sample_offsets, lay.frame_offsets = so, fo
    lay.total_samples, lay.total_frames = so[-1], fo[-1]
    lay.c_sample_offsets, lay.c_frame_offsets = al._lib.i64_array(so), al._lib.i64_array(fo)
    dev = torch.from_numpy(np.concatenate(wavs)).cuda()
    mfcc, mel, pdb = (x.cpu().numpy() for x in al.frontend_device(plan, dev, lay))
    ref = al.calc_MFCC_input_batch(wavs, **HP)
    for u, (o, t) in enumerate(zip(fo, lay.frames)):
        for got, want, name in zip((mfcc[o:o + t], mel[o:o + t], pdb[o:o + t]), ref[u], ("MFCC", "M_dB", "P_dB")):
            # same arithmetic up to the DCT sums: float64 in the scalar pass B, centred float32 chains in the hp pass B
            assert_close(got, want, rtol=1e-6, atol=5e-6 if name == "MFCC" else 2e-7, what=f"utt {u}/{name}")


@pytest.mark.parametrize("n_chunks,n_streams,ramp", [(8, 3, True), (3, 2, False), (64, 4, True), (1, 1, True)])
def test_host_pipeline_equals_batch_call(al, n_chunks, n_streams, ramp):
    """FrontendPipeline (bench.py's e2e path: pinned H2D -> kernels -> pinned D2H over several streams, ramped chunk
    sizes) must return exactly what one ragged batch call returns, for every chunking."""
    lens = [16000, 8001, 24000, 4497, 12345, 400, 31999, 16000, 7777, 20000, 9600, 480]
    wavs = [synth.utterance(300 + i, n / 16000.0)[:n] for i, n in enumerate(lens)]
    kw = dict(sr=16000, n_fft=400, win_length=400, hop_length=80, n_mels=80, n_mfcc=40, window="hann", pre_emphasis=0.97,
              mfcc_normaleze_first_mfcc=True, mfcc_norm_factor=0.01, calc_mfcc_derivate=True, M_dB_norm_factor=0.01,
              P_dB_norm_factor=0.01, mean_abs_amp_norm=0.003, clip_output=True)
    pipe = al.FrontendPipeline(lens, n_chunks=n_chunks, n_streams=n_streams, ramp=ramp, **kw)
    assert sum(b - a for a, b, *_ in pipe.chunks) == len(lens)
    pipe.load(wavs)
    pipe.run()
    pipe.run()                                              # buffers are reused: a second pass must give the same
    want = al.calc_MFCC_input_batch(wavs, **HP)
    for got, ref in zip(pipe.views(), want):
        for g, r in zip(got, ref):
            np.testing.assert_array_equal(g, r)


def test_frontend_gain_is_numpy_exact_inside_the_hot_path(al):
    """The gain sc_frontend_batch applies is float32(m / float64(np.abs(y).mean())) bit for bit: with every other stage
    switched to identity-like settings the outputs of two utterances that differ only by an exact power-of-two scale
    are bit-identical, and a 1-ulp change of one sample's magnitude that flips the float32 mean flips the output."""
    y = synth.utterance(41, 1.0)
    a = al.calc_MFCC_input(y, **HP)
    b = al.calc_MFCC_input((y * np.float32(4.0)).astype(np.float32), **HP)      # mean scales exactly, gain / 4 exactly
    for p, q in zip(a, b):
        np.testing.assert_array_equal(p, q)


def test_same_layout_reuses_tables_and_results_do_not_change(al):
    """Second call with the same ragged layout skips the descriptor upload and k_fe_setup (layout cache): one launch
    fewer, same bits; a different layout in between invalidates it."""
    import torch
    lens = [48000, 16001, 7999, 64000]
    wavs = [synth.utterance(50 + i, n / 16000.0)[:n] for i, n in enumerate(lens)]
    plan = al.DspPlan(**{**dict(sr=16000, n_fft=400, win_length=400, hop_length=80, n_mels=80, n_mfcc=40),
                         **{k: HP[k] for k in ("pre_emphasis", "mfcc_norm_factor", "M_dB_norm_factor", "P_dB_norm_factor",
                                               "mean_abs_amp_norm", "clip_output", "calc_mfcc_derivate",
                                               "mfcc_normaleze_first_mfcc")}})
    lay = al.FrontendLayout(lens, 80)
    dev = torch.zeros(lay.total_samples, dtype=torch.float32, device="cuda")
    for w, o in zip(wavs, lay.sample_offsets):
        dev[o:o + len(w)] = torch.from_numpy(w).cuda()
    def same(x, y):          # rows between utterances (16-byte alignment padding) are never written: compare real rows
        return all(torch.equal(p[o:o + t], q[o:o + t]) for p, q in zip(x, y) for o, t in zip(lay.frame_offsets, lay.frames))
    al.launch_count_reset()
    first = [t.clone() for t in al.frontend_device(plan, dev, lay)]
    n1 = al.launch_count()
    second = al.frontend_device(plan, dev, lay)
    n2 = al.launch_count() - n1
    assert n2 == n1 - 1
    assert same(first, second)
    lay2 = al.FrontendLayout(lens[:2], 80)
    al.frontend_device(plan, dev, lay2)
    third = al.frontend_device(plan, dev, lay)
    assert same(first, third)


def test_no_allocation_after_reserve(al):
    """SURVEY.md section 8(b), last row: after sc_plan_reserve, compute calls within the bounds never allocate."""
    import torch
    from speech_cloner_b200 import _lib
    lib = _lib.load()
    lens = [64000] * 6 + [30001, 555]
    plan = al.DspPlan(sr=16000, n_fft=400, win_length=400, hop_length=80, n_mels=80, n_mfcc=40, calc_mfcc_derivate=True)
    lay = al.FrontendLayout(lens, 80)
    plan.reserve(lay.total_samples, lay.total_frames, len(lens))
    dev = torch.rand(lay.total_samples, dtype=torch.float32, device="cuda") - 0.5
    out = al.frontend_device(plan, dev, lay)                   # allocates torch outputs only
    glay = al._GlLayout([t for t in lay.frames], 80)
    amp = torch.rand((glay.frame_offsets[-1], 201), dtype=torch.float32, device="cuda")
    ph = torch.rand_like(amp)
    wav = torch.empty(glay.sample_offsets[-1], dtype=torch.float32, device="cuda")
    out64 = torch.empty(glay.sample_offsets[-1], dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()
    n0 = lib.sc_alloc_count()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(2):
        al.frontend_device(plan, dev, lay, out)
        _lib.check(lib.sc_power_to_amp_batch(plan._h, amp.data_ptr(), glay.c_frame_offsets, glay.c_frame_counts, len(lens),
                                             0.01, 1.2, amp.data_ptr(), st), "p2a")
        al.griffin_lim_device(plan, amp, ph, glay, 3, None, wav)
        _lib.check(lib.sc_deemph_renorm_batch(plan._h, wav.data_ptr(), glay.c_sample_offsets, glay.c_sample_lengths, len(lens),
                                              0.97, 0.045, out64.data_ptr(), st), "deemph")
    torch.cuda.synchronize()
    assert lib.sc_alloc_count() == n0
    assert torch.cuda.mem_get_info()[0] == free0                # cudaMemGetInfo delta = 0


def test_plan_is_single_threaded_and_ordered_across_streams(al):
    """A second thread entering a call on the same plan is refused; the public API hands every thread its own plan;
    a call on another stream is ordered after the previous one."""
    import threading
    import torch
    lens = [64000] * 64
    plan = al.DspPlan(sr=16000, n_fft=400, win_length=400, hop_length=80, n_mels=80, n_mfcc=40)
    lay = al.FrontendLayout(lens, 80)
    dev = torch.rand(lay.total_samples, dtype=torch.float32, device="cuda") - 0.5
    ref = [t.clone() for t in al.frontend_device(plan, dev, lay)]
    errors, oks = [], []

    def worker():
        for _ in range(30):
            try:
                al.frontend_device(plan, dev, lay)
                oks.append(1)
            except ValueError as e:
                errors.append(str(e))
    ths = [threading.Thread(target=worker) for _ in range(4)]
    [t.start() for t in ths]; [t.join() for t in ths]
    assert oks and all("single-threaded" in e for e in errors)
    torch.cuda.synchronize()
    # other stream: results identical, no explicit synchronisation by the caller
    s2 = torch.cuda.Stream()
    out_a = al.frontend_device(plan, dev, lay)
    with torch.cuda.stream(s2):
        out_b = al.frontend_device(plan, dev, lay)
    torch.cuda.synchronize()
    for p, q, r in zip(ref, out_a, out_b):
        for o, t in zip(lay.frame_offsets, lay.frames):
            assert torch.equal(p[o:o + t], q[o:o + t]) and torch.equal(p[o:o + t], r[o:o + t])
    # the reference-signature functions are thread-safe: the plan cache is keyed by thread
    y = synth.utterance(7, 0.5)
    want = al.calc_MFCC_input(y, **HP)
    res = []

    def api_worker():
        res.append(al.calc_MFCC_input(y, **HP))
    ths = [threading.Thread(target=api_worker) for _ in range(4)]
    [t.start() for t in ths]; [t.join() for t in ths]
    assert len(res) == 4
    for r in res:
        for p, q in zip(want, r):
            np.testing.assert_array_equal(p, q)


def test_all_zero_utterance_raises_like_librosa(al):
    import torch
    with pytest.raises(ValueError):
        al.calc_MFCC_input(np.zeros(8000, dtype=np.float32), **HP)
    with pytest.raises(ValueError):
        al.calc_MFCC_input(torch.zeros(8000, dtype=torch.float32, device="cuda"), **HP)
    # the flag is cleared by the poll: the next good call works
    al.calc_MFCC_input(torch.from_numpy(synth.utterance(3, 0.5)).cuda(), **HP)
    hp1 = dict(HP); hp1["mean_abs_amp_norm"] = 1.0           # no gain requested: zeros are legal input (:125)
    al.calc_MFCC_input(np.zeros(8000, dtype=np.float32), **hp1)
