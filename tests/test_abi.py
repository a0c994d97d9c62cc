"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol of include/speechdsp.h."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "speechdsp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    lib = C.CDLL(built_lib)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in speechdsp.h but not exported"


def test_python_prototypes_cover_header(built_lib):
    from speech_cloner_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == _declared()
    _lib.load()


def test_params_struct_layout_matches_c(built_lib, tmp_path):
    from speech_cloner_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "speechdsp.h"\n'
                   'int main(){printf("%zu %zu %zu %zu", sizeof(sc_params), offsetof(sc_params, fft_precision),'
                   ' offsetof(sc_params, pre_emphasis), offsetof(sc_params, window_host));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    size, o_prec, o_pre, o_win = map(int, subprocess.check_output([str(exe)]).split())
    P = _lib.ScParams
    assert (C.sizeof(P), P.fft_precision.offset, P.pre_emphasis.offset, P.window_host.offset) == (size, o_prec, o_pre, o_win)


def test_sass_is_sm100a(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu(built_lib):
    """Without a CUDA device the library must refuse to compute (SC_ERR_NO_DEVICE), never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from speech_cloner_b200 import _lib, audio_lib
    lib = _lib.load()
    p = _lib.ScParams()
    p.sample_rate, p.n_fft, p.win_length, p.hop_length, p.n_mels, p.n_mfcc = 16000, 400, 400, 80, 80, 40
    h = C.c_void_p()
    assert lib.sc_plan_create(C.byref(p), C.byref(h)) == _lib.SC_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.sc_last_error()
    import numpy as np
    with pytest.raises(_lib.SpeechDspError):
        audio_lib.calc_MFCC_input(np.zeros(1000, dtype=np.float32) + 0.1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "speech_cloner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_window_gather_argument_checks_need_no_gpu(built_lib):
    """sc_window_gather validates its host arguments before any CUDA call: bad calls are SC_ERR_INVALID with a message,
    an empty batch is a no-op (sound_ds.py:262-350 never yields an empty batch, but the ABI must not launch one)."""
    from speech_cloner_b200 import _lib
    lib = _lib.load()
    P = C.c_void_p
    src, dst = (P * 1)(0x1000), (P * 1)(0x2000)
    w1, w0 = (C.c_int64 * 1)(80), (C.c_int64 * 1)(0)
    dev = P(0x3000)

    def call(s=src, d=dst, w=w1, n_arrays=1, rows=10, first=dev, valid=dev, n_windows=0, n_t=4):
        return lib.sc_window_gather(s, d, w, n_arrays, rows, first, valid, n_windows, n_t, None)
    assert call() == _lib.SC_OK                                       # zero windows: nothing is launched
    for bad in (dict(n_arrays=0), dict(n_arrays=5), dict(w=w0), dict(n_t=0), dict(n_windows=-1), dict(rows=-1),
                dict(first=None), dict(valid=None), dict(s=(P * 1)(0)), dict(d=(P * 1)(0x2002))):
        assert call(**bad) == _lib.SC_ERR_INVALID, bad
        assert b"sc_window_gather" in lib.sc_last_error()
