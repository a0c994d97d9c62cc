"""B200-native audio-DSP hot path of socom20/speech-cloner (front-end + Griffin-Lim).

``speech_cloner_b200.audio_lib`` mirrors the reference's ``audio_lib`` module; the arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of ``include/speechdsp.h``.
"""
from . import audio_lib  # noqa: F401
from .audio_lib import (calc_MFCC_input, calc_MFCC_input_batch, calc_PHN_target, calc_inv_preemphasis,  # noqa: F401
                        calc_preemphasis, from_power_to_wav, from_power_to_wav_batch, griffin_lim_alg,
                        griffin_lim_batch)

__version__ = "0.1.0"
