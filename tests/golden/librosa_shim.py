"""The slice of the librosa-0.6 API that /root/reference/audio_lib.py calls, so that the REFERENCE FILE ITSELF can be
imported and run in this container (librosa, matplotlib and sounddevice are not installable here).

TEST INFRASTRUCTURE ONLY (used by make_reference_vectors.py).  With this shim in ``sys.modules`` every line of
audio_lib.py:12-308 executes unmodified - the gain, pre-/de-emphasis, dtype chain, normalisation, deltas, clipping, the
Griffin-Lim loop with its NumPy random phase, the ``realse`` power law - and only the librosa primitives below are
substituted.  The large ones (stft, istft, filters.mel, filters.dct) are the oracle's restatements, which
tests/test_oracle_pins.py pins against torch.stft / torch.istft / transformers / torchaudio / scipy; the small ones are
restated here line by line from librosa 0.6.3 (core/spectrum.py), independently of the oracle's folded versions.
"""
import sys
import types

import numpy as np

from oracle import audio_lib_oracle as _o


def stft(y, n_fft=2048, hop_length=None, win_length=None, window='hann', center=True, dtype=np.complex64,
         pad_mode='reflect'):
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    assert center and pad_mode == 'reflect' and dtype == np.complex64, "only the reference's call pattern"
    return _o.stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window)


def istft(stft_matrix, hop_length=None, win_length=None, window='hann', center=True, dtype=np.float32, length=None):
    n_fft = 2 * (stft_matrix.shape[0] - 1)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    assert center and dtype == np.float32 and length is None, "only the reference's call pattern"
    return _o.istft(stft_matrix, hop_length=hop_length, win_length=win_length, window=window)


def magphase(D, power=1):
    mag = np.abs(D)
    mag **= power
    phase = np.exp(1.j * np.angle(D))
    return mag, phase


def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
    S = np.asarray(S)
    if amin <= 0:
        raise ValueError('amin must be strictly positive')
    magnitude = np.abs(S)
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        if top_db < 0:
            raise ValueError('top_db must be non-negative')
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def amplitude_to_db(S, ref=1.0, amin=1e-5, top_db=80.0):
    S = np.asarray(S)
    magnitude = np.abs(S)
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    power = np.square(magnitude, out=magnitude)
    return power_to_db(power, ref=ref_value ** 2, amin=amin ** 2, top_db=top_db)


def db_to_power(S_db, ref=1.0):
    return ref * np.power(10.0, 0.1 * S_db)


def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm=1):
    assert not htk and norm == 1, "only the reference's call pattern"
    return _o.mel_filterbank(sr, n_fft, n_mels, fmin=fmin, fmax=fmax)


def dct(n_filters, n_input):
    return _o.dct_basis(n_filters, n_input)


def install():
    """Put ``librosa`` (+ core, filters, display), ``matplotlib.pyplot`` stand-ins into sys.modules; returns the names
    added so that the caller can remove them again."""
    added = []

    def module(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        added.append(name)
        return m

    prim = dict(stft=stft, istft=istft, magphase=magphase, power_to_db=power_to_db, amplitude_to_db=amplitude_to_db,
                db_to_power=db_to_power)
    core = module("librosa.core", **prim)
    filters = module("librosa.filters", mel=mel, dct=dct)
    display = module("librosa.display")
    module("librosa", core=core, filters=filters, display=display, __version__="0.6.3-shim", **prim)
    plt = module("matplotlib.pyplot")
    module("matplotlib", pyplot=plt)
    return added


def uninstall(added):
    for name in added:
        sys.modules.pop(name, None)
