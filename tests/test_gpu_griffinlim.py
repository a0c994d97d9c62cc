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
