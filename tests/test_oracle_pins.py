"""Pin the CPU oracle piece by piece against independent libraries (no GPU).

The reference ships no tests / vectors and librosa is not installable here, so these are the
available cross-checks: torch.stft, transformers.audio_utils (a librosa
port), scipy.fft.dct, torchaudio, plus analytic identities.
"""
import numpy as np
import pytest
import scipy.fft
import torch

from oracle import audio_lib_oracle as o
from speech_cloner_b200 import synth


@pytest.fixture(scope="module")
def y():
    return synth.utterance(123, 1.5)


@pytest.mark.parametrize("n_fft,hop,win", [(400, 80, 400), (512, 128, 512), (400, 40, 320)])
def test_stft_matches_torch(y, n_fft, hop, win):
    got = o.stft(y.astype(np.float64), n_fft, hop, win)
    w = torch.hann_window(win, periodic=True, dtype=torch.float64)
    ref = torch.stft(torch.from_numpy(y.astype(np.float64)), n_fft, hop, win, w, center=True, pad_mode="reflect",
                     return_complex=True).numpy()
    assert got.shape == ref.shape == (1 + n_fft // 2, 1 + len(y) // hop)
    assert got.dtype == np.complex64
    assert np.abs(got - ref).max() <= 1e-7 * np.abs(ref).max() + 1e-12


def test_reflect_index_matches_numpy_pad():
    for n, pad in [(5, 3), (2, 7), (10, 25), (1, 4)]:
        a = np.arange(n, dtype=np.float64) + 1
        want = np.pad(a, pad, mode="reflect")
        got = a[o.reflect_index(np.arange(-pad, n + pad), n)]
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("sr,n_fft,n_mels", [(16000, 400, 80), (16000, 400, 128), (8000, 256, 40), (22050, 2048, 128)])
def test_mel_matches_transformers_and_torchaudio(sr, n_fft, n_mels):
    from transformers.audio_utils import mel_filter_bank
    import torchaudio
    M = o.mel_filterbank(sr, n_fft, n_mels)
    ref = mel_filter_bank(1 + n_fft // 2, n_mels, 0.0, sr / 2, sr, norm="slaney", mel_scale="slaney").T
    assert np.abs(M - ref).max() < 1e-12
    ta = torchaudio.functional.melscale_fbanks(1 + n_fft // 2, 0.0, sr / 2, n_mels, sr, norm="slaney",
                                               mel_scale="slaney").numpy().T
    assert np.abs(M - ta).max() < 1e-6
    assert (np.count_nonzero(M, axis=0) <= 2).all()          # the structure the CUDA mel kernel relies on


def test_dct_matches_scipy():
    x = np.random.default_rng(0).standard_normal((80, 9))
    assert np.abs(o.dct_basis(40, 80) @ x - scipy.fft.dct(x, type=2, norm="ortho", axis=0)[:40]).max() < 1e-13


def test_db_conversions_match_transformers():
    from transformers.audio_utils import amplitude_to_db, power_to_db
    S = np.abs(np.random.default_rng(1).standard_normal((201, 50))) ** 2 * 1e-3
    got = o.power_to_db(S.astype(np.float32))
    ref = power_to_db(S.astype(np.float32), reference=1.0, min_value=1e-10, db_range=80.0)
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-5)
    got = o.amplitude_to_db(S)
    ref = amplitude_to_db(S, reference=1.0, min_value=1e-5, db_range=80.0)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-10)


def test_istft_inverts_stft_and_matches_torch(y):
    F = o.stft(y.astype(np.float64), 400, 80)
    w = o.istft(F, 80)
    assert w.dtype == np.float32 and w.shape == (80 * (F.shape[1] - 1),)
    assert np.abs(w - y[: w.size]).max() < 5e-7
    ref = torch.istft(torch.from_numpy(F.astype(np.complex128)), 400, 80, 400,
                      torch.hann_window(400, periodic=True, dtype=torch.float64), center=True).numpy()
    assert np.abs(w - ref[: w.size]).max() < 5e-7


def test_window_sumsquare_values():
    wss = o.window_sumsquare("hann", 30, 80, 400, 400)
    assert wss.dtype == np.float32 and wss.shape == (400 + 80 * 29,)
    assert wss[200] == np.float32(1.4375)                      # first kept sample (SURVEY.md §7 "Edge semantics")
    np.testing.assert_allclose(wss[400:2300], 1.875, rtol=2e-7)
    w2 = o.padded_window("hann", 400, 400) ** 2
    np.testing.assert_allclose(wss[:80], w2[:80].astype(np.float32), rtol=1e-7)


def test_preemphasis_pair_matches_lfilter(y):
    from scipy import signal
    np.testing.assert_allclose(o.calc_preemphasis(y, 0.97), signal.lfilter([1, -0.97], [1], y), rtol=0, atol=1e-15)
    np.testing.assert_allclose(o.calc_inv_preemphasis(y, 0.97), signal.lfilter([1], [1, -0.97], y), rtol=0, atol=0)
    rt = o.calc_inv_preemphasis(o.calc_preemphasis(y, 0.97), 0.97)
    np.testing.assert_allclose(rt, y, atol=1e-12)


def test_frontend_shapes_dtypes_and_ranges(y):
    mfcc, mel, pdb = o.calc_MFCC_input(y, **synth.HP_ENC)
    T = 1 + len(y) // 80
    assert mfcc.shape == (T, 80) and mel.shape == (T, 80) and pdb.shape == (T, 201)
    assert mfcc.dtype == mel.dtype == pdb.dtype == np.float32
    assert pdb.min() == 0.0 and 0.0 < pdb.max() <= 0.8 + 1e-6            # top_db = 80 times 0.01
    assert mel.min() == 0.0 and (mfcc[0, 40:] == 0).all() and (mfcc[-1, 40:] == 0).all()
    assert mfcc[0, 0] == 0.0                                             # c0 shift (:221)
    np.testing.assert_allclose(mfcc[1:-1, 40:], np.clip(2 * (mfcc[2:, :40].astype(np.float64) - mfcc[:-2, :40]), -1, 1),
                               atol=2e-7)


def test_frontend_equals_composition_of_pieces(y):
    """calc_MFCC_input restated from its parts with torch.stft as the FFT (independent of oracle.stft)."""
    hp = dict(synth.HP_ENC)
    g = np.float32(np.float64(0.003) / np.float64(np.abs(y).mean()))
    ys = y * g
    pe = np.concatenate([[ys[0]], ys[1:].astype(np.float64) - 0.97 * ys[:-1].astype(np.float64)])
    F = torch.stft(torch.from_numpy(pe), 400, 80, 400, torch.hann_window(400, periodic=True, dtype=torch.float64),
                   center=True, pad_mode="reflect", return_complex=True).numpy().astype(np.complex64)
    P = np.abs(F) ** 2
    pdb = 10 * np.log10(np.maximum(1e-10, P)); pdb = np.maximum(pdb, pdb.max() - 80)
    want_p = np.clip(0.01 * (pdb - pdb.min()), -1, 1).T
    got = o.calc_MFCC_input(y, **hp)
    np.testing.assert_allclose(got[2], want_p, atol=3e-6)
    M = o.mel_filterbank(16000, 400, 80) @ P.astype(np.float64)
    mdb = 10 * np.log10(np.maximum(1e-10, M ** 2)); mdb = np.maximum(mdb, mdb.max() - 80)
    np.testing.assert_allclose(got[1], np.clip(0.01 * (mdb - mdb.min()), -1, 1).T, atol=3e-6)
    cc = scipy.fft.dct(mdb, type=2, norm="ortho", axis=0)[:40].T
    cc[:, 0] -= cc[0, 0]
    np.testing.assert_allclose(got[0][:, :40], np.clip(0.01 * cc, -1, 1), atol=3e-6)


def test_griffin_lim_converges_and_is_seed_reproducible(y):
    P = o.calc_MFCC_input(y, **synth.HP_ENC)[2][:120]
    amp = np.sqrt(np.power(np.float32(10.0), np.float32(0.1) * (P.T / np.float32(0.01) - np.float32(80.0))))
    np.random.seed(5)
    ph = np.pi * np.random.rand(*amp.shape)
    log = []
    w1 = o.griffin_lim_alg(amp, 400, 80, num_iters=30, verbose=False, phase0=ph, rms_log=log)
    np.random.seed(5)
    w2 = o.griffin_lim_alg(amp, 400, 80, num_iters=30, verbose=False)
    np.testing.assert_array_equal(w1, w2)                   # phase0=None draws the same numbers (:255)
    assert w1.dtype == np.float32 and w1.shape == (80 * 119,)
    assert len(log) == 29 and log[-1] < 0.2 * log[0]        # iterates settle
    sc = lambda w: np.linalg.norm(np.abs(o.stft(w, 400, 80)) - amp) / np.linalg.norm(amp)
    assert sc(w1) < sc(o.griffin_lim_alg(amp, 400, 80, num_iters=2, verbose=False, phase0=ph))


def test_from_power_to_wav_contract(y):
    P = o.calc_MFCC_input(y, **synth.HP_ENC)[2][:60]
    np.random.seed(1)
    w = o.from_power_to_wav(P, hop_length=80, win_length=400, mean_abs_amp_norm=0.045, n_iter=5, realse=1.2, verbose=False)
    assert w.dtype == np.float64 and w.shape == (80 * 59,)
    np.testing.assert_allclose(np.abs(w).mean(), 0.045, rtol=1e-12)


def test_phn_target_matches_reference_logic():
    """calc_PHN_target restated vs a literal transcription of the reference loop (audio_lib.py:51-85)."""
    rng = np.random.default_rng(3)
    bounds = np.sort(rng.choice(np.arange(100, 15900), size=11, replace=False))
    edges = np.concatenate([[0], bounds, [16000]])
    phn_v = [(int(a), int(b), f"p{i % 5}") for i, (a, b) in enumerate(zip(edges[:-1], edges[1:]))]
    conv = {f"p{i}": i for i in range(5)}
    yy = np.zeros(16000, dtype=np.float32)

    def literal(y, phn_v, d, hop_length=40, win_length=400):
        n = int(y.shape[0] / hop_length) + 1
        half = win_length // 2
        out, i_phn = [], 0
        for i_s in range(n):
            s, e = i_s * hop_length - half, i_s * hop_length + win_length - half
            while phn_v[i_phn][1] <= s and i_phn + 1 < len(phn_v):
                i_phn += 1
            a = min(phn_v[i_phn][1], e) - max(phn_v[i_phn][0], s)
            if i_phn + 1 < len(phn_v):
                b = min(phn_v[i_phn + 1][1], e) - max(phn_v[i_phn + 1][0], s)
                out.append(d[phn_v[i_phn][2]] if a >= b else d[phn_v[i_phn + 1][2]])
            else:
                out.append(d[phn_v[i_phn][2]])
        return np.array(out, dtype=np.int32)

    for hop, win in [(80, 400), (40, 400), (128, 512)]:
        want = literal(yy, phn_v, conv, hop, win)
        np.testing.assert_array_equal(o.calc_PHN_target(yy, phn_v, conv, hop, win), want)
        from speech_cloner_b200.audio_lib import calc_PHN_target
        np.testing.assert_array_equal(calc_PHN_target(yy, phn_v, conv, hop, win), want)
