"""GPU parity of Griffin-Lim (/root/reference/audio_lib.py:249-308) against the oracle.

Bar (BASELINE.json north_star): same fixed initial phase, SNR >= 40 dB against the oracle waveform and
spectral convergence |STFT(y)| vs A within 1e-3 of the oracle's.
"""
import numpy as np
import pytest

from oracle import audio_lib_oracle as oracle
from speech_cloner_b200 import synth
from tests.util import snr_db

pytestmark = pytest.mark.gpu

HP = dict(synth.HP_ENC)
GL = dict(P_dB_norm_factor=0.01, pre_emphasis=0.97, hop_length=80, win_length=400, mean_abs_amp_norm=0.045,
          n_fft=None)


@pytest.fixture(scope="module")
def al(built_lib):
    from speech_cloner_b200 import audio_lib
    return audio_lib


def _pdb(seed, seconds):
    return oracle.calc_MFCC_input(synth.utterance(seed, seconds), **HP)[2]


def _phase0(seed, shape):
    np.random.seed(seed)
    return np.pi * np.random.rand(*shape)


def _spectral_convergence(y, amp, hop=80, win=400):
    X = np.abs(oracle.stft(np.asarray(y, dtype=np.float32), n_fft=win, hop_length=hop, win_length=win))
    return float(np.linalg.norm(X - amp) / np.linalg.norm(amp))


@pytest.mark.parametrize("n_iter", [1, 2, 25])
def test_griffin_lim_alg_matches_oracle(al, n_iter):
    P = _pdb(3000, 1.0)[:160]
    amp = np.sqrt(np.power(np.float32(10.0), np.float32(0.1) * (P.T / np.float32(0.01) - np.float32(80.0))))
    ph = _phase0(3000, amp.shape)
    want = oracle.griffin_lim_alg(amp, 400, 80, num_iters=n_iter, verbose=False, phase0=ph)
    got = al.griffin_lim_alg(amp, 400, 80, num_iters=n_iter, verbose=False, phase0=ph)
    assert got.dtype == np.float32 and got.shape == want.shape == (80 * 159,)
    assert snr_db(got, want) >= 60.0
    assert abs(_spectral_convergence(got, amp) - _spectral_convergence(want, amp)) < 1e-3


def test_rms_delta_log_matches_reference_print(al, capsys):
    P = _pdb(3001, 0.6)[:100]
    amp = np.sqrt(np.power(np.float32(10.0), np.float32(0.1) * (P.T / np.float32(0.01) - np.float32(80.0))))
    ph = _phase0(1, amp.shape)
    log = []
    oracle.griffin_lim_alg(amp, 400, 80, num_iters=6, verbose=False, phase0=ph, rms_log=log)
    al.griffin_lim_alg(amp, 400, 80, num_iters=6, verbose=True, phase0=ph)
    lines = [l for l in capsys.readouterr().out.splitlines() if "mrse_delta" in l]
    assert len(lines) == 5
    got = [float(l.split("=")[-1]) for l in lines]
    np.testing.assert_allclose(got, log, rtol=2e-3)


@pytest.mark.parametrize("realse", [1.0, 1.2])
def test_from_power_to_wav_matches_oracle(al, realse):
    """Config-3-shaped call (test.py:148-168) at a size the oracle finishes in seconds."""
    P = _pdb(3002, 1.5)[:240]
    ph = _phase0(3002, (201, 240))
    want = oracle.from_power_to_wav(P, n_iter=40, realse=realse, verbose=False, phase0=ph, **GL)
    got = al.from_power_to_wav(P, n_iter=40, realse=realse, verbose=False, phase0=ph, **GL)
    assert got.dtype == np.float64 and got.shape == want.shape
    assert snr_db(got, want) >= 40.0
    np.testing.assert_allclose(np.abs(got).mean(), 0.045, rtol=1e-9)


def test_batch_equals_single(al):
    Ps = [_pdb(3100 + i, 1.0)[:n] for i, n in enumerate([64, 150, 33, 2, 9])]
    phs = [_phase0(i, (201, P.shape[0])) for i, P in enumerate(Ps)]
    got = al.from_power_to_wav_batch(Ps, n_iter=12, phase0s=phs, **GL)
    for P, ph, g in zip(Ps, phs, got):
        single = al.from_power_to_wav(P, n_iter=12, verbose=False, phase0=ph, **GL)
        np.testing.assert_array_equal(g, single)
        want = oracle.from_power_to_wav(P, n_iter=12, verbose=False, phase0=ph, **GL)
        assert snr_db(g, want) >= 40.0


def test_no_deemphasis_path(al):
    P = _pdb(3200, 0.8)[:100]
    ph = _phase0(5, (201, 100))
    kw = dict(GL); kw["pre_emphasis"] = 0
    want = oracle.from_power_to_wav(P, n_iter=8, verbose=False, phase0=ph, **kw)
    got = al.from_power_to_wav(P, n_iter=8, verbose=False, phase0=ph, **kw)
    assert snr_db(got, want) >= 40.0


@pytest.mark.parametrize("geom", [dict(hop_length=40, win_length=800),        # from_power_to_wav's defaults (:281-282)
                                  dict(hop_length=128, win_length=512)])
def test_generic_geometry(al, geom):
    hop, win = geom["hop_length"], geom["win_length"]
    y = synth.utterance(3300, 1.0)
    kw = dict(HP); kw.update(hop_length=hop, win_length=win, n_fft=None)
    P = oracle.calc_MFCC_input(y, **kw)[2][:60]
    ph = _phase0(11, (win // 2 + 1, 60))
    a = dict(GL); a.update(geom)
    want = oracle.from_power_to_wav(P, n_iter=6, verbose=False, phase0=ph, **a)
    got = al.from_power_to_wav(P, n_iter=6, verbose=False, phase0=ph, **a)
    assert got.shape == want.shape
    assert snr_db(got, want) >= 40.0


def test_random_phase_default_uses_numpy_global_state(al):
    P = _pdb(3400, 0.5)[:50]
    np.random.seed(123)
    a = al.from_power_to_wav(P, n_iter=3, verbose=False, **GL)
    b = al.from_power_to_wav(P, n_iter=3, verbose=False, phase0=_phase0(123, (201, 50)), **GL)
    np.testing.assert_array_equal(a, b)


def test_chunk_step_equals_unchunked(al):
    """Time-chunked projection: two chunks with halos reproduce the whole-signal step."""
    import ctypes as C
    import torch
    from speech_cloner_b200 import _lib
    from speech_cloner_b200.audio_lib import DspPlan, _GlLayout, griffin_lim_device
    lib = _lib.load()
    plan = DspPlan.get(n_fft=400, win_length=400, hop_length=80)
    T = 301
    P = _pdb(3500, 2.0)[:T]
    amp = torch.from_numpy(np.ascontiguousarray(
        np.sqrt(np.power(np.float32(10.0), np.float32(0.1) * (P / np.float32(0.01) - np.float32(80.0)))))).cuda()
    ph = torch.from_numpy(_phase0(9, (T, 201)).astype(np.float32)).cuda()
    lay = _GlLayout([T], 80)
    whole = griffin_lim_device(plan, amp, ph, lay, 3)[: 80 * (T - 1)].clone()

    Lw = 80 * (T - 1)
    cuts = [0, 80 * 140, Lw]                                   # two ranks
    st = torch.cuda.current_stream().cuda_stream
    state = torch.zeros(Lw, dtype=torch.float32, device="cuda")
    nxt = torch.zeros_like(state)
    for it in range(3):
        for r in range(2):
            lo, hi = cuts[r], cuts[r + 1]
            f_lo = max(0, lo // 80 - 4); f_hi = min(T, hi // 80 + 6)
            w_lo = max(0, lo - 1000); w_hi = min(Lw, hi + 1000)
            rc = lib.sc_griffinlim_chunk_step(
                plan._h, amp[f_lo:f_hi].contiguous().data_ptr(),
                ph[f_lo:f_hi].contiguous().data_ptr() if it == 0 else None,
                f_lo, f_hi - f_lo, T, state[w_lo:w_hi].contiguous().data_ptr() if it else None, w_lo, w_hi - w_lo,
                nxt[lo:hi].data_ptr(), lo, hi - lo, st)
            _lib.check(rc, "chunk_step")
            torch.cuda.synchronize()
        state, nxt = nxt, state
    np.testing.assert_array_equal(state.cpu().numpy(), whole.cpu().numpy())


# ------------------------------------------------------------------------------------------------------------------
# Time-chunked long-form path on ONE GPU: the ranks of a 2- / 3- / 4-GPU run are emulated one after the other (the halo
# exchanges are tensor copies), so the driver's single-GPU test run covers sc_griffinlim_chunk_run, the windowed
# de-emphasis carry and the canonical sums.  tests/test_gpu_distributed.py repeats it over NCCL.
def _emulate(world, T, n_iter, k, realse=1.0, hop=80, n_fft=400, pdb_seed=3600):
    import torch
    from speech_cloner_b200 import distributed as D
    from speech_cloner_b200.audio_lib import DspPlan
    plan = DspPlan.get(n_fft=n_fft, win_length=n_fft, hop_length=hop)
    kw = dict(HP); kw.update(hop_length=hop, win_length=n_fft, n_fft=None)
    base = oracle.calc_MFCC_input(synth.utterance(pdb_seed, 3.0), **kw)[2]
    P = np.concatenate([base] * (T // base.shape[0] + 1))[:T]
    ph = _phase0(17, (n_fft // 2 + 1, T))
    ranks = [D.ChunkedGriffinLim(T, hop, n_fft, steps_per_exchange=k, plan=plan, world=world, rank=r) for r in range(world)]
    P_dev = torch.from_numpy(np.ascontiguousarray(P)).cuda()
    ph_dev = torch.from_numpy(np.ascontiguousarray(ph.T).astype(np.float32)).cuda()
    return plan, P, ph, ranks, P_dev, ph_dev


def _emulated_exchange(ranks, bufs, e_los, width):
    """What ChunkedGriffinLim._exchange does over NCCL, as copies between the emulated ranks' buffers."""
    for r, cg in enumerate(ranks):
        if cg.hi <= cg.lo:
            continue
        for nb in (r - 1, r + 1):
            if not cg._live(nb):
                continue
            o = ranks[nb]
            if nb < r:      # my left halo <- the left neighbour's last `width` samples
                bufs[r][cg.lo - e_los[r] - width: cg.lo - e_los[r]] = bufs[nb][o.hi - e_los[nb] - width: o.hi - e_los[nb]]
            else:
                w = min(width, cg.total - cg.hi)
                bufs[r][cg.hi - e_los[r]: cg.hi - e_los[r] + w] = bufs[nb][o.lo - e_los[nb]: o.lo - e_los[nb] + w]


@pytest.mark.parametrize("world,T,n_iter,k,realse", [(2, 449, 7, 3, 1.0), (3, 1200, 9, 20, 1.0), (4, 2001, 12, 4, 1.2),
                                                     (2, 321, 4, 2, 1.0)])
def test_chunk_run_emulated_ranks_bit_identical(al, world, T, n_iter, k, realse):
    """Communication-avoiding chunked from_power_to_wav == the single-GPU call, bit for bit."""
    import torch
    from speech_cloner_b200 import _lib
    from speech_cloner_b200 import distributed as D
    geom = dict(hop=40, n_fft=800) if T == 321 else {}
    plan, P, ph, ranks, P_dev, ph_dev = _emulate(world, T, n_iter, k, realse, **geom)
    hop, n_fft = ranks[0].hop, ranks[0].n_fft
    kw = dict(GL); kw.update(hop_length=hop, win_length=n_fft)
    whole = al.from_power_to_wav(P, n_iter=n_iter, realse=realse, verbose=False, phase0=ph, **kw)
    lib, st = plan._lib, torch.cuda.current_stream().cuda_stream
    assert all(cg.hi > cg.lo for cg in ranks)
    # ---- prologue: block partials of the rows every rank owns, "all_gather" = concatenation in rank order
    amps = []
    parts = []
    for cg in ranks:
        f_lo, f_hi = cg.frame_range(n_iters=n_iter)
        o_lo, o_hi = cg.own_frame_range()
        part = torch.zeros(2 * (-(-(o_hi - o_lo) // cg.align)), dtype=torch.float64, device="cuda")
        if realse != 1.0:
            _lib.check(lib.sc_p2a_chunk_partial(plan._h, P_dev[o_lo:].data_ptr(), o_hi - o_lo, realse, part.data_ptr(), st), "partial")
        parts.append(part)
    allp = torch.cat(parts).contiguous()
    for cg in ranks:
        f_lo, f_hi = cg.frame_range(n_iters=n_iter)
        amp = torch.empty((f_hi - f_lo, n_fft // 2 + 1), dtype=torch.float32, device="cuda")
        _lib.check(lib.sc_p2a_chunk_apply(plan._h, P_dev[f_lo:f_hi].contiguous().data_ptr(), f_hi - f_lo, 0.01, realse,
                                          allp.data_ptr() if realse != 1.0 else None, allp.shape[0] // 2, amp.data_ptr(), st), "apply")
        amps.append(amp)
    # ---- iterations in rounds of k with emulated exchanges
    a, b, e_los = [], [], []
    for cg in ranks:
        e_lo, e_hi = cg.ext_range(n_iters=n_iter)
        a.append(torch.zeros(e_hi - e_lo, dtype=torch.float32, device="cuda")); b.append(torch.zeros_like(a[-1])); e_los.append(e_lo)
    kk = ranks[0]._k_for(n_iter)
    done = 0
    while done < n_iter:
        n = min(kk, n_iter - done)
        for r, cg in enumerate(ranks):
            f_lo, f_hi = cg.frame_range(n_iters=n_iter)
            e_lo, e_hi = cg.ext_range(n_iters=n_iter)
            cg._round(amps[r], ph_dev[f_lo:f_hi].contiguous() if done == 0 else None, f_lo, f_hi, a[r], b[r], e_lo, e_hi, n)
        if n & 1:
            a, b = b, a
        done += n
        if done < n_iter:
            _emulated_exchange(ranks, a, e_los, min(kk, n_iter - done) * ranks[0].halo)
    chunks = [a[r][cg.lo - e_los[r]: cg.hi - e_los[r]].contiguous() for r, cg in enumerate(ranks)]
    # float32 Griffin-Lim state: compare against the whole-signal kernel through the public single call
    plan1 = plan
    from speech_cloner_b200.audio_lib import _GlLayout, griffin_lim_device
    amp_whole = torch.empty_like(P_dev)
    lay = _GlLayout([T], hop)
    _lib.check(lib.sc_power_to_amp_batch(plan1._h, P_dev.data_ptr(), lay.c_frame_offsets, lay.c_frame_counts, 1, 0.01, realse,
                                         amp_whole.data_ptr(), st), "p2a")
    gl_whole = griffin_lim_device(plan1, amp_whole, ph_dev, lay, n_iter)[: hop * (T - 1)]
    assert torch.equal(torch.cat(chunks), gl_whole)
    # ---- epilogue: local responses, the left neighbour's last `win`, apply, gathered block sums, renorm
    win = int(lib.sc_deemph_chunk_window(0.97))
    assert win == 12
    locs = []
    for r, cg in enumerate(ranks):
        n_chunks = -(-(cg.hi - cg.lo) // 256)
        loc = torch.zeros(win + n_chunks, dtype=torch.float64, device="cuda")
        _lib.check(lib.sc_deemph_chunk_local(plan._h, chunks[r].data_ptr(), cg.lo, cg.hi - cg.lo, cg.total, 0.97, loc[win:].data_ptr(), st), "local")
        locs.append(loc)
    for r in range(1, world):
        locs[r][:win] = locs[r - 1][-win:]
    outs, sums = [], []
    for r, cg in enumerate(ranks):
        out = torch.empty(cg.hi - cg.lo, dtype=torch.float64, device="cuda")
        sm = torch.zeros(-(-(cg.hi - cg.lo) // cg.sum_block), dtype=torch.float64, device="cuda")
        _lib.check(lib.sc_deemph_chunk_apply(plan._h, chunks[r].data_ptr(), cg.lo, cg.hi - cg.lo, cg.total, 0.97, locs[r].data_ptr(), win,
                                             out.data_ptr(), sm.data_ptr(), st), "apply")
        outs.append(out); sums.append(sm)
    alls = torch.cat(sums).contiguous()
    for r, cg in enumerate(ranks):
        _lib.check(lib.sc_renorm_chunk(plan._h, outs[r].data_ptr(), cg.hi - cg.lo, alls.data_ptr(), alls.shape[0], cg.total, 0.045, st), "renorm")
    got = torch.cat(outs).cpu().numpy()
    np.testing.assert_array_equal(got, whole)                           # float64 epilogue: bit-identical too


def test_deemphasis_windowed_carry_matches_lfilter(al):
    """The truncated carry series of the chunked scan against scipy's sequential filter, long signal, float64 accuracy."""
    from scipy import signal
    rng = np.random.default_rng(5)
    x = (0.05 * rng.standard_normal(300001)).astype(np.float32)
    got = al.calc_inv_preemphasis(x, 0.97)
    want = signal.lfilter([1.0], [1.0, -0.97], x.astype(np.float64))
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-13)
    # a coefficient too close to 1 for a short window takes the sequential fallback
    got2 = al.calc_inv_preemphasis(x[:70000], 0.9999)
    want2 = signal.lfilter([1.0], [1.0, -0.9999], x[:70000].astype(np.float64))
    np.testing.assert_allclose(got2, want2, rtol=1e-10, atol=1e-11)
    # float64 input is not rounded to float32 first (scipy keeps float64)
    x64 = x.astype(np.float64) + 1e-11
    np.testing.assert_allclose(al.calc_inv_preemphasis(x64, 0.97), signal.lfilter([1.0], [1.0, -0.97], x64), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(al.calc_preemphasis(x64, 0.97), signal.lfilter([1.0, -0.97], [1.0], x64), rtol=0, atol=1e-15)


def test_zero_iterations_follow_the_reference(al):
    P = _pdb(3400, 0.5)[:50]
    amp = np.ones((201, 50), dtype=np.float32)
    assert al.griffin_lim_alg(amp, 400, 80, num_iters=0, verbose=False) is None        # audio_lib.py:252 `wav = None`
    with pytest.raises(ValueError):
        al.from_power_to_wav(P, n_iter=0, verbose=False, **GL)


def test_non_finite_result_raises_like_the_reference(al):
    """An all-zero map with realse != 1 is 0 / 0 at audio_lib.py:296: the reference's librosa.stft then rejects the NaN
    waveform from the second iteration on (found by scripts/soak.py); with one iteration no stft runs and NaN comes back."""
    P = np.zeros((40, 201), dtype=np.float32)
    kw = dict(GL); kw.update(verbose=False, realse=1.2, phase0=_phase0(5, (201, 40)))
    with pytest.raises(ValueError):
        al.from_power_to_wav(P, n_iter=3, **kw)
    with pytest.raises(ValueError), np.errstate(all="ignore"):
        oracle.from_power_to_wav(P, n_iter=3, **kw)
    with np.errstate(all="ignore"):
        assert np.isnan(oracle.from_power_to_wav(P, n_iter=1, **kw)).all()
    assert np.isnan(al.from_power_to_wav(P, n_iter=1, **kw)).all()
    kw["realse"] = 1.0                                            # no power law: amplitude 1e-4 everywhere, finite
    assert np.isfinite(al.from_power_to_wav(P, n_iter=3, **kw)).all()
