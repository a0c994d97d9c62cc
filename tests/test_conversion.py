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
