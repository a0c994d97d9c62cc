"""Full-size checks at BASELINE.json's configurations (GPU).  Where the oracle would take minutes, the
checks are size-independent properties of the reference's definition instead of element-wise parity."""
import numpy as np
import pytest

from oracle import audio_lib_oracle as oracle
from speech_cloner_b200 import synth
from tests.util import assert_close, snr_db

pytestmark = pytest.mark.gpu
HP = dict(synth.HP_ENC)


@pytest.fixture(scope="module")
def al(built_lib):
    from speech_cloner_b200 import audio_lib
    return audio_lib


def test_config2_256x4s_properties(al):
    """configs[1]: 256 utterances x 4 s.  Properties that hold for any input by construction of
    audio_lib.py:89-244, plus oracle parity on a sample of the batch."""
    base = synth.batch(2, 16, 4.0)
    wavs = [np.ldexp(base[i % 16], i // 16 - 4) for i in range(256)]      # 16 utterances at 16 power-of-two levels
    feats = al.calc_MFCC_input_batch(wavs, **HP)
    assert len(feats) == 256
    for i, (mfcc, mel, pdb) in enumerate(feats):
        assert mfcc.shape == (801, 80) and mel.shape == (801, 80) and pdb.shape == (801, 201)
        assert pdb.min() == 0.0 and pdb.max() <= 0.8 + 1e-6               # min shift, top_db = 80 dB * 0.01
        assert mel.min() == 0.0 and mel.max() <= 0.8 + 1e-6
        assert mfcc[0, 0] == 0.0                                           # c0 shift (:221)
        assert not mfcc[0, 40:].any() and not mfcc[-1, 40:].any()          # zero first / last delta row (:227)
        d = np.clip(2.0 * (mfcc[2:, :40].astype(np.float64) - mfcc[:-2, :40]), -1, 1)
        np.testing.assert_allclose(mfcc[1:-1, 40:], d, atol=3e-7)
        assert np.abs(mfcc).max() <= 1.0
    # the mean-|y| gain (:126) makes the features invariant to the input level.  Power-of-two levels scale
    # every float32 operation of the gain stage exactly, so the invariance must hold bit for bit
    # (for other gains the reference itself moves near-floor bins by ~1e-4: its float32 samples re-round).
    for i in range(16):
        for k in range(1, 16):
            for a, b in zip(feats[i], feats[i + 16 * k]):
                np.testing.assert_array_equal(a, b, err_msg=f"level invariance {i}/{k}")
    for i in (0, 5, 255):
        want = oracle.calc_MFCC_input(wavs[i], **HP)
        for a, b in zip(feats[i], want):
            assert_close(a, b, what=f"utt {i}")


def test_config3_one_spectrogram_200_iterations(al):
    """configs[2] shape, full length: (1000, 201) decoder-shaped spectrogram, 200 iterations, fixed phase0,
    the arguments of test.py:148-156.  SNR >= 40 dB and spectral convergence within 1e-3 of the oracle's."""
    P = oracle.calc_MFCC_input(synth.utterance(3000, 5.0), **HP)[2][:1000]
    np.random.seed(3000)
    ph = np.pi * np.random.rand(201, 1000)
    kw = dict(P_dB_norm_factor=0.01, pre_emphasis=0.97, hop_length=80, win_length=400, mean_abs_amp_norm=0.045,
              n_iter=200, n_fft=None, realse=1.0, verbose=False, phase0=ph)
    want = oracle.from_power_to_wav(P, **kw)
    got = al.from_power_to_wav(P, **kw)
    assert got.shape == want.shape == (79920,)
    assert snr_db(got, want) >= 40.0
    amp = np.sqrt(np.power(np.float32(10.0), np.float32(0.1) * (P.T / np.float32(0.01) - np.float32(80.0))))

    def sc(y):
        pe = oracle.calc_preemphasis(y, 0.97).astype(np.float32)          # back to the pre-emphasised domain
        X = np.abs(oracle.stft(pe, 400, 80))
        s = np.vdot(X, amp) / np.vdot(X, X)                                # the final renorm rescales y
        return float(np.linalg.norm(s * X - amp) / np.linalg.norm(amp))
    assert abs(sc(got) - sc(want)) < 1e-3


def test_config3_batch_64_is_consistent(al):
    """64 spectrograms in one ragged batch give the same waveforms as single calls (independence of jobs)."""
    base = [oracle.calc_MFCC_input(synth.utterance(3100 + i, 5.0), **HP)[2][:1000] for i in range(4)]
    Ps = [base[i % 4] for i in range(64)]
    phs = []
    for i in range(64):
        np.random.seed(3000 + i % 4)
        phs.append(np.pi * np.random.rand(201, 1000))
    kw = dict(P_dB_norm_factor=0.01, pre_emphasis=0.97, hop_length=80, win_length=400, mean_abs_amp_norm=0.045,
              n_iter=20, realse=1.0)
    out = al.from_power_to_wav_batch(Ps, phase0s=phs, **kw)
    for i in range(4, 64):
        np.testing.assert_array_equal(out[i], out[i % 4])
    single = al.from_power_to_wav(Ps[1], verbose=False, phase0=phs[1], n_fft=None, **kw)
    np.testing.assert_array_equal(out[1], single)


def test_config3_batch_64_against_the_oracle_at_200_iterations(al):
    """configs[2] exactly: 64 spectrograms x 5 s in ONE batch call (host arrays, pipelined staging), 200 iterations,
    fixed initial phases; four of them (spread over the batch, so every staging group is covered) are checked against the
    oracle at the full 200 iterations: SNR >= 40 dB (the batch used to be self-consistency only)."""
    wavs = synth.batch(3, 64, 5.0)
    Ps = [oracle.calc_MFCC_input(w, **HP)[2][:1000] for w in wavs[:8]]
    Ps = [Ps[i % 8] for i in range(64)]
    phs = []
    for i in range(64):
        np.random.seed(3000 + i)
        phs.append(np.pi * np.random.rand(201, 1000))
    kw = dict(P_dB_norm_factor=0.01, pre_emphasis=0.97, hop_length=80, win_length=400, mean_abs_amp_norm=0.045,
              n_iter=200, realse=1.0)
    out = al.from_power_to_wav_batch(Ps, phase0s=phs, **kw)
    assert len(out) == 64 and all(o.shape == (79920,) and o.dtype == np.float64 for o in out)
    for i in (0, 21, 42, 63):
        want = oracle.from_power_to_wav(Ps[i], n_fft=None, verbose=False, phase0=phs[i], **kw)
        assert snr_db(out[i], want) >= 40.0, i
        np.testing.assert_allclose(np.abs(out[i]).mean(), 0.045, rtol=1e-9)


def test_config5_shaped_sweep_is_consistent(al):
    """configs[4] in the shape bench.py times it (16 distinct TIMIT-shaped 3 s + 16 ARCTIC-shaped 4 s utterances tiled), at a
    tenth of the 10 h: every replica of an utterance must give the same bits wherever it sits in the ragged batch (tile
    tables, packed offsets beyond 2^31 bytes are 64-bit), and the distinct ones must match the oracle."""
    import torch
    timit = synth.batch(5, 16, 3.0, ds_norm=(0.0, 10.0))
    arctic = synth.batch(6, 16, 4.0)
    wavs = [timit[i % 16] for i in range(600)] + [arctic[i % 16] for i in range(450)]
    dev = [torch.from_numpy(w).cuda() for w in timit + arctic]
    feats = al.calc_MFCC_input_batch([dev[i % 16] for i in range(600)] + [dev[16 + i % 16] for i in range(450)], **HP)
    assert len(feats) == 1050
    for i, f in enumerate(feats):
        first = feats[i % 16] if i < 600 else feats[600 + (i - 600) % 16]
        for a, b in zip(f, first):
            assert a.shape == b.shape and torch.equal(a, b), f"replica {i} differs from its first copy"
    for i in (0, 7, 600, 611, 1049):
        want = oracle.calc_MFCC_input(wavs[i], **HP)
        for a, b in zip(feats[i], want):
            assert_close(a.cpu().numpy(), b, what=f"utt {i}")
