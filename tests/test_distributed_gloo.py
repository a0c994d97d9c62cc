"""World-size-2 `gloo` tests (CPU) of the multi-GPU host logic: sharding, ragged gather, halo exchange.

The per-rank compute is injected (the oracle), so only the distribution logic of
speech_cloner_b200.distributed is under test here; the CUDA kernels are covered by the -m gpu tests.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import audio_lib_oracle as oracle
from speech_cloner_b200 import distributed as D
from speech_cloner_b200 import synth

HP = dict(synth.HP_ENC)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(fn, world, *args):
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn, args), nprocs=world, join=True)


def _entry(rank, world, port, fn, args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def test_shard_by_frames_is_balanced_and_deterministic():
    rng = np.random.default_rng(0)
    lens = rng.integers(8000, 128000, size=101).tolist()
    for w in (1, 2, 4, 8):
        sh = D.shard_by_frames(lens, w, 80)
        assert sorted(i for s in sh for i in s) == list(range(101))
        load = [sum(1 + lens[i] // 80 for i in s) for s in sh]
        assert max(load) - min(load) <= max(1 + n // 80 for n in lens)
        assert sh == D.shard_by_frames(lens, w, 80)
    assert D.shard_by_frames([100, 200], 4, 80)[2:] == [[], []]


def test_chunk_bounds_cover_the_signal():
    for T, w in [(240001, 8), (1001, 2), (301, 2), (100, 4), (30, 8)]:
        b = D.chunk_bounds(T, w, 80)
        assert b[0][0] == 0 and b[-1][1] == 80 * (T - 1)
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert all(lo % (112 * 80) == 0 for lo, hi in b if hi > lo)       # tile grid (28 hops) and 256-sample scan grid
    lo8 = [hi - lo for lo, hi in D.chunk_bounds(240001, 8, 80)]
    assert max(lo8) <= 1.01 * min(lo8)                                       # balanced at chapter length


def test_chunk_geometry_is_derived_from_the_stft_geometry():
    """halo and alignment follow (n_fft, hop), not the (400, 80) constants."""
    assert D.chunk_geometry(400, 80) == (112, 480, 6, 8960)
    align, halo, halo_frames, blk = D.chunk_geometry(800, 40)           # from_power_to_wav's signature defaults
    assert halo_frames == 800 // 40 + 1 and halo == halo_frames * 40 and halo >= 800 + 40
    assert (align * 40) % 256 == 0 and align % 20 == 0 and blk == align * 40
    cg = D.ChunkedGriffinLim(4001, 40, 800, step=lambda *a: None, world=4, rank=1, steps_per_exchange=3)
    lo, hi = cg.bounds[1]
    assert cg.ext_range(n_iters=10) == (lo - 3 * halo, hi + 3 * halo)
    assert cg.frame_range(n_iters=10) == (lo // 40 - 3 * halo_frames, hi // 40 + 3 * halo_frames + 1)
    assert cg.frame_range(n_iters=2) == (lo // 40 - 2 * halo_frames, hi // 40 + 2 * halo_frames + 1)
    with pytest.raises(ValueError):
        D.ChunkedGriffinLim(171, 40, 800, step=lambda *a: None, world=4, rank=0)     # second chunk: 400 samples < one halo


def _oracle_batch(wavs, **hp):
    return [oracle.calc_MFCC_input(w, **hp) for w in wavs]


def _run_featurize(rank, world):
    wavs = [synth.utterance(500 + i, s) for i, s in enumerate([0.5, 0.3, 0.8, 0.2, 0.4])]
    got = D.featurize_sharded(wavs, compute=_oracle_batch, gather=True, **HP)
    assert len(got) == len(wavs)
    for w, g in zip(wavs, got):
        want = oracle.calc_MFCC_input(w, **HP)
        for a, b in zip(g, want):
            np.testing.assert_array_equal(a.numpy(), b)
    local = D.featurize_sharded(wavs, compute=_oracle_batch, gather=False, **HP)
    assert sorted(local) == D.shard_by_frames([len(w) for w in wavs], world, 80)[rank]


def test_featurize_sharded_gathers_ragged_outputs():
    _spawn(_run_featurize, 2)


def _run_gather_packed(rank, world):
    sizes = [5, 3]
    mine = torch.arange(sizes[rank], dtype=torch.float64) + 10 * rank
    parts = D.gather_packed(mine, sizes)
    assert [p.tolist() for p in parts] == [[0, 1, 2, 3, 4], [10, 11, 12]]


def test_gather_packed_ragged():
    _spawn(_run_gather_packed, 2)


def _oracle_step_factory(T, n_fft=400, hop=80):
    """Chunk step restated with the oracle: zero-extend what the rank knows, run one whole-signal projection."""
    total = hop * (T - 1)
    bins = n_fft // 2 + 1

    def step(amp, phase0, first_frame, n_local, n_total, wav_in, wav_first, wav_count, wav_out, out_first, out_count):
        A = np.zeros((bins, T), dtype=np.float32)
        A[:, first_frame:first_frame + n_local] = amp.numpy().T
        if phase0 is not None:
            P = np.zeros((bins, T))
            P[:, first_frame:first_frame + n_local] = phase0.numpy().T.astype(np.float64)
            S = A * np.exp(1j * P)
        else:
            y = np.zeros(total, dtype=np.float32)
            y[wav_first:wav_first + wav_count] = wav_in.numpy()
            X = oracle.stft(y, n_fft, hop)
            S = A * np.exp(np.complex64(1j) * np.angle(X))
        out = oracle.istft(S, hop, n_fft)
        wav_out.copy_(torch.from_numpy(out[out_first:out_first + out_count].copy()))
    return step


def _run_chunked_gl(rank, world, T, n_fft, hop, k, n_iter):
    bins = n_fft // 2 + 1
    kw = dict(HP); kw.update(hop_length=hop, win_length=n_fft, n_fft=None)
    y = np.concatenate([synth.utterance(77 + i, 1.0) for i in range(-(-(hop * T) // 16000))])
    P = oracle.calc_MFCC_input(y, **kw)[2][:T]
    assert P.shape == (T, bins)
    amp = np.sqrt(np.power(np.float32(10.0), np.float32(0.1) * (P / np.float32(0.01) - np.float32(80.0)))).astype(np.float32)
    np.random.seed(3)
    ph = (np.pi * np.random.rand(T, bins)).astype(np.float32)
    whole = oracle.griffin_lim_alg(amp.T, n_fft, hop, num_iters=n_iter, verbose=False, phase0=ph.T.astype(np.float64))
    gl = D.ChunkedGriffinLim(T, hop, n_fft, step=_oracle_step_factory(T, n_fft, hop), steps_per_exchange=k)
    assert all(b > a for a, b in gl.bounds), "both ranks must own samples in this test"
    f_lo, f_hi = gl.frame_range(n_iters=n_iter)
    chunk = gl.run(torch.from_numpy(amp[f_lo:f_hi]), torch.from_numpy(ph[f_lo:f_hi]), n_iter)
    assert chunk.shape[0] == gl.hi - gl.lo
    ref = whole[gl.lo:gl.hi]
    err = np.abs(chunk.numpy() - ref).max()
    assert err <= 1e-5 * np.abs(whole).max() + 1e-9, err        # same projection: the k-step halo is sufficient
    full = gl.gather(chunk, dst=0)
    if rank == 0:
        assert full.shape[0] == hop * (T - 1)
        np.testing.assert_allclose(full.numpy(), whole, atol=1e-5 * np.abs(whole).max() + 1e-9)
    else:
        assert full is None


@pytest.mark.parametrize("T,n_fft,hop,k,n_iter", [
    (337, 400, 80, 1, 3),        # exchange every iteration (the round-1 schedule)
    (337, 400, 80, 2, 5),        # communication-avoiding: rounds of 2, 2, 1 iterations
    (337, 400, 80, 20, 4),       # one round: no exchange at all after the widened first strip
    (321, 800, 40, 2, 3),        # from_power_to_wav's default geometry (generic kernels): halo 840 samples, 21 frames
])
def test_chunked_griffin_lim_halo_exchange(T, n_fft, hop, k, n_iter):
    _spawn(_run_chunked_gl, 2, T, n_fft, hop, k, n_iter)
