"""Host-side logic of the Python mirror that needs no GPU: layouts, validation order, plan keys."""
import numpy as np
import pytest

from speech_cloner_b200 import audio_lib as al
from speech_cloner_b200 import synth


def test_frontend_layout_alignment_and_counts():
    lens = [48000, 16001, 7999, 81, 80, 79, 1]
    lay = al.FrontendLayout(lens, 80)
    assert lay.frames == [1 + n // 80 for n in lens] == [601, 201, 100, 2, 2, 1, 1]
    assert all(o % 4 == 0 for o in lay.sample_offsets + lay.frame_offsets)        # 16-byte aligned rows
    for i, n in enumerate(lens):
        assert lay.sample_offsets[i + 1] - lay.sample_offsets[i] >= n
        assert lay.frame_offsets[i + 1] - lay.frame_offsets[i] >= lay.frames[i]
    assert list(lay.c_sample_lengths) == lens


def test_gl_layout():
    lay = al._GlLayout([1000, 2, 301], 80)
    assert lay.samples == [79920, 80, 24000]
    assert all(o % 4 == 0 for o in lay.sample_offsets + lay.frame_offsets)


def test_validation_happens_before_any_device_work():
    hp = dict(synth.HP_ENC)
    for bad in ([0.0, 1.0], np.zeros(10, dtype=np.int32), np.zeros((2, 8), dtype=np.float32),
                np.array([], dtype=np.float32), np.array([0.0, np.inf], dtype=np.float32)):
        with pytest.raises(ValueError):
            al.calc_MFCC_input(bad, **hp)


def test_window_array_matches_scipy():
    from scipy import signal
    np.testing.assert_array_equal(al._window_array("hann", 400), signal.get_window("hann", 400, fftbins=True))
    np.testing.assert_array_equal(al._window_array("hamm", 400), signal.get_window("hamming", 400, fftbins=True))
    np.testing.assert_array_equal(al._window_array(("kaiser", 4.0), 64), signal.get_window(("kaiser", 4.0), 64))
    with pytest.raises(ValueError):
        al._window_array(np.ones(5), 400)


def test_signatures_match_reference():
    """Positional order and defaults of the five reference functions (audio_lib.py:12, :31, :89-104, :249, :278-287)."""
    import inspect
    sig = lambda f: [(p.name, p.default) for p in inspect.signature(f).parameters.values()]
    E = inspect.Parameter.empty
    assert sig(al.calc_preemphasis) == [("wav", E), ("coeff", 0.97)]
    assert sig(al.calc_inv_preemphasis) == [("preem_wav", E), ("coeff", 0.97)]
    assert sig(al.calc_PHN_target) == [("y", E), ("phn_v", E), ("phn_conv_d", E), ("hop_length", 40), ("win_length", 400)]
    assert sig(al.calc_MFCC_input) == [
        ("y", E), ("sr", 16000), ("pre_emphasis", 0.97), ("hop_length", 40), ("win_length", 400), ("n_mels", 128),
        ("n_mfcc", 40), ("n_fft", None), ("window", "hann"), ("mfcc_normaleze_first_mfcc", True),
        ("mfcc_norm_factor", 0.01), ("calc_mfcc_derivate", False), ("M_dB_norm_factor", 0.01),
        ("P_dB_norm_factor", 0.01), ("mean_abs_amp_norm", 0.003), ("clip_output", True)]
    assert sig(al.griffin_lim_alg)[:6] == [("stft_amp", E), ("win_length", E), ("hop_length", E), ("num_iters", 300),
                                           ("n_fft", None), ("verbose", True)]
    assert sig(al.from_power_to_wav)[:10] == [
        ("P", E), ("P_dB_norm_factor", 0.01), ("pre_emphasis", 0.97), ("hop_length", 40), ("win_length", 800),
        ("mean_abs_amp_norm", 0.01), ("n_iter", 200), ("n_fft", None), ("realse", 1.0), ("verbose", True)]


def test_synth_is_seeded_and_shaped():
    a, b = synth.utterance(7, 1.0), synth.utterance(7, 1.0)
    np.testing.assert_array_equal(a, b)
    assert a.dtype == np.float32 and a.shape == (16000,) and 0.05 < np.abs(a).max() < 0.2
    assert np.abs(synth.utterance(7, 1.0, ds_norm=(0.0, 10.0)) - 10 * a).max() < 1e-6


def test_signatures_match_the_reference_source():
    """Pinned to the reference's own source: parse /root/reference/audio_lib.py with ast (it cannot be imported here:
    librosa / matplotlib are absent) and compare names, order and defaults of the five functions.  Skipped on the GPU
    box, where /root/reference does not exist; tests/golden/reference_signatures.json is the travelling copy."""
    import ast
    import inspect
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    gold = os.path.join(here, "golden", "reference_signatures.json")
    names = ["calc_preemphasis", "calc_inv_preemphasis", "calc_PHN_target", "calc_MFCC_input", "griffin_lim_alg",
             "from_power_to_wav"]

    def from_source(path):
        tree = ast.parse(open(path).read())
        out = {}
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name in names:
                args = node.args.args
                defaults = [None] * (len(args) - len(node.args.defaults)) + list(node.args.defaults)
                out[node.name] = [[a.arg, "<required>" if d is None else repr(ast.literal_eval(d))] for a, d in zip(args, defaults)]
        return out
    ref_path = "/root/reference/audio_lib.py"
    if os.path.exists(ref_path):
        parsed = from_source(ref_path)
        assert sorted(parsed) == sorted(names)
        assert parsed == json.load(open(gold)), "tests/golden/reference_signatures.json is stale: rerun tests/golden/make_signatures.py"
    want = json.load(open(gold))
    for name in names:
        sig = inspect.signature(getattr(al, name))
        got = [[p.name, "<required>" if p.default is inspect.Parameter.empty else repr(p.default)] for p in sig.parameters.values()]
        assert got[:len(want[name])] == want[name], name
        extra = got[len(want[name]):]
        assert all(e[0] in ("phase0",) for e in extra), f"{name}: only the documented phase0 keyword may be added, got {extra}"


def test_pairwise_subtree_bound_of_the_abs_kernel():
    """The |y| kernel stages one sub-tree of NumPy's pairwise summation per CTA (csrc/fe_kernels.cuh).  NumPy splits a node
    at n/2 rounded DOWN to a multiple of 8, so the sub-trees at depth abs_depth(n) hold up to kAbsSubtree + 15 samples,
    not kAbsSubtree - the bound the staging buffer (kAbsSubtreeMax = kAbsSubtree + 16) and the seven slot levels rely on.
    The constants are read from the source so that the test follows the kernel."""
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "speech_cloner_b200", "csrc",
                            "fe_kernels.cuh")).read()
    K = int(re.search(r"#define SC_ABS_SUBTREE (\d+)", src).group(1))
    slack = int(re.search(r"kAbsSubtreeMax = kAbsSubtree \+ (\d+)", src).group(1))

    def depth(n):
        d = 0
        while (n >> d) > K:
            d += 1
        return d

    def split(sizes):
        out = []
        for x in sizes:
            n2 = x // 2
            n2 -= n2 % 8
            out += [n2, x - n2]
        return out

    def subtrees(n):
        s = [n]
        for _ in range(depth(n)):
            s = split(s)
        return s

    rng = np.random.default_rng(0)
    lengths = {K * (1 << D) + k for D in range(9) for k in range(-40, 3)} | {int(x) for x in rng.integers(1, 1 << 21, size=4000)}
    worst = max(max(subtrees(n)) for n in lengths if n > 0)
    assert K < worst <= K + slack - 1                    # the bound is real (8 015 at K = 8 000) and inside the buffer
    # seven more levels bring every sub-tree of up to K + slack samples down to leaves NumPy sums without splitting
    for s in list(range(K - 40, K + slack + 1)) + [int(x) for x in rng.integers(129, K, size=300)]:
        leaves = [s]
        for _ in range(7):
            leaves = [y for x in leaves for y in (split([x]) if x > 128 else [x])]
        assert max(leaves) <= 128
