"""compound stitching and write_wav normalisation: host path against the literal
transcription of test.py:46-84 in the oracle; the CUDA-tensor path must give the same rows."""
import numpy as np
import pytest

from oracle import audio_lib_oracle as oracle


@pytest.mark.parametrize("n0,n1,T", [(1, 0, 8), (2, 1, 8), (5, 4, 400), (4, 3, 10), (3, 2, 7), (3, 5, 12), (6, 2, 16), (3, 2, 3)])
def test_compound_equals_the_reference_loop(n0, n1, T):
    from speech_cloner_b200 import conversion as cv
    rng = np.random.default_rng(n0 * 100 + n1 * 10 + T)
    y0, y1 = rng.standard_normal((n0, T, 5)).astype(np.float32), rng.standard_normal((n1, T, 5)).astype(np.float32)
    want = oracle.compound(y0, y1)
    got = cv.compound(y0, y1)
    assert got.shape == want.shape and (got == want).all()


@pytest.mark.parametrize("T,n_times,t_s,t_e", [(2450, 400, 1, 60), (2400, 400, 0, 60), (900, 400, 0, 3), (1300, 400, 1, 60), (4000, 200, 2, 10)])
def test_window_batches_equal_the_reference_slicing(T, n_times, t_s, t_e):
    """test.py:92-128: zero padding to a whole number of windows, [t_s, t_e) cut, aligned and half-offset batches."""
    from speech_cloner_b200 import conversion as cv
    rng = np.random.default_rng(T + n_times)
    mfcc, mel, stft = (rng.standard_normal((T, w)).astype(np.float32) for w in (80, 80, 201))
    cfg = dict(hop_length=80, n_timesteps=n_times, sample_rate=16000)
    w0, w1, mel_t, stft_t, n_s, n_e = oracle.window_batches(mfcc, mel, stft, cfg, t_s, t_e)
    got = cv.window_batches(mfcc, mel, stft, cfg, t_s, t_e)
    assert (got["n_s"], got["n_e"]) == (n_s, n_e)
    np.testing.assert_array_equal(got["mfcc_input0"], w0)
    assert (got["mfcc_input1"] is None) == (w1 is None)
    if w1 is not None:
        np.testing.assert_array_equal(got["mfcc_input1"], w1)
        # an identity "decoder" stitched back gives the original frames (compound keeps 3/4 + middles + 3/4)
        np.testing.assert_array_equal(cv.stitch_predictions(got["mfcc_input0"], got["mfcc_input1"]), oracle.compound(w0, w1))
    np.testing.assert_array_equal(got["mel_true"], mel_t)
    np.testing.assert_array_equal(got["stft_true"], stft_t)


def test_window_batches_raise_like_the_reference():
    from speech_cloner_b200 import conversion as cv
    x = np.zeros((300, 4), np.float32)
    with pytest.raises(Exception):
        cv.window_batches(x, x, x, dict(hop_length=80, n_timesteps=400, sample_rate=16000), t_s=5, t_e=60)


def test_normalize_wav():
    from speech_cloner_b200 import conversion as cv
    y = np.array([0.1, -0.4, 0.2], np.float64)
    assert np.allclose(cv.normalize_wav(y), oracle.normalize_wav(y)) and np.max(np.abs(cv.normalize_wav(y))) == 1.0
    z = np.zeros(4, np.float32)
    assert (cv.normalize_wav(z) == 0).all()


@pytest.mark.gpu
def test_compound_on_device_and_render(built_lib):
    import torch
    from speech_cloner_b200 import audio_lib as al, conversion as cv
    rng = np.random.default_rng(3)
    y0 = (0.8 * rng.random((3, 40, 201))).astype(np.float32)
    y1 = (0.8 * rng.random((2, 40, 201))).astype(np.float32)
    got = cv.compound(torch.from_numpy(y0).cuda(), torch.from_numpy(y1).cuda())
    want = oracle.compound(y0, y1)
    # the batching in front of the decoder, on device tensors: same rows as the reference slicing, no host round trip
    rng2 = np.random.default_rng(9)
    mfcc, mel, stft = (rng2.standard_normal((1300, w)).astype(np.float32) for w in (80, 80, 201))
    cfg = dict(hop_length=80, n_timesteps=400, sample_rate=16000)
    wb = cv.window_batches(*(torch.from_numpy(a).cuda() for a in (mfcc, mel, stft)), cfg, 1, 60)
    w0, w1, mel_t, stft_t, n_s, n_e = oracle.window_batches(mfcc, mel, stft, cfg, 1, 60)
    assert wb["mfcc_input0"].is_cuda and (wb["mfcc_input0"].cpu().numpy() == w0).all()
    assert (wb["mfcc_input1"].cpu().numpy() == w1).all() and (wb["stft_true"].cpu().numpy() == stft_t).all()
    assert (cv.stitch_predictions(wb["mfcc_input0"], wb["mfcc_input1"]).cpu().numpy() == oracle.compound(w0, w1)).all()
    assert got.is_cuda and tuple(got.shape) == want.shape and (got.cpu().numpy() == want).all()
    np.random.seed(1)
    ph = np.pi * np.random.rand(201, want.shape[0])
    kw = dict(P_dB_norm_factor=0.01, pre_emphasis=0.97, hop_length=80, win_length=400, mean_abs_amp_norm=0.045,
              n_iter=5, verbose=False, phase0=ph)
    y_dev = cv.render_windows(torch.from_numpy(y0).cuda(), torch.from_numpy(y1).cuda(), **kw)
    y_ref = oracle.normalize_wav(oracle.from_power_to_wav(want, **kw))
    assert y_dev.is_cuda
    y = y_dev.cpu().numpy()
    snr = 10 * np.log10(np.sum(y_ref ** 2) / np.sum((y - y_ref) ** 2))
    assert snr >= 40.0 and abs(np.max(np.abs(y)) - 1.0) < 1e-12
