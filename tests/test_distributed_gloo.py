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
        assert all(lo % (28 * 80) == 0 for lo, hi in b if hi > lo)


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


def _oracle_step_factory(T):
    """Chunk step restated with the oracle: zero-extend what the rank knows, run one whole-signal projection."""
    total = 80 * (T - 1)

    def step(amp, phase0, first_frame, n_local, n_total, wav_in, wav_first, wav_count, wav_out, out_first, out_count):
        A = np.zeros((201, T), dtype=np.float32)
        A[:, first_frame:first_frame + n_local] = amp.numpy().T
        if phase0 is not None:
            P = np.zeros((201, T))
            P[:, first_frame:first_frame + n_local] = phase0.numpy().T.astype(np.float64)
            S = A * np.exp(1j * P)
        else:
            y = np.zeros(total, dtype=np.float32)
            y[wav_first:wav_first + wav_count] = wav_in.numpy()
            X = oracle.stft(y, 400, 80)
            S = A * np.exp(np.complex64(1j) * np.angle(X))
        out = oracle.istft(S, 80, 400)
        wav_out.copy_(torch.from_numpy(out[out_first:out_first + out_count].copy()))
    return step


def _run_chunked_gl(rank, world):
    T = 141                                                   # 11 200 samples, 28-hop cut at 6 720
    P = oracle.calc_MFCC_input(synth.utterance(77, 1.0), **HP)[2][:T]
    amp = np.sqrt(np.power(np.float32(10.0), np.float32(0.1) * (P / np.float32(0.01) - np.float32(80.0)))).astype(np.float32)
    np.random.seed(3)
    ph = (np.pi * np.random.rand(T, 201)).astype(np.float32)
    n_iter = 4
    whole = oracle.griffin_lim_alg(amp.T, 400, 80, num_iters=n_iter, verbose=False, phase0=ph.T.astype(np.float64))
    gl = D.ChunkedGriffinLim(T, 80, 400, step=_oracle_step_factory(T))
    f_lo, f_hi = gl.frame_range()
    chunk = gl.run(torch.from_numpy(amp[f_lo:f_hi]), torch.from_numpy(ph[f_lo:f_hi]), n_iter)
    assert chunk.shape[0] == gl.hi - gl.lo
    ref = whole[gl.lo:gl.hi]
    err = np.abs(chunk.numpy() - ref).max()
    assert err <= 1e-5 * np.abs(whole).max() + 1e-9, err        # same projection, halo is sufficient
    full = gl.gather(chunk, dst=0)
    if rank == 0:
        assert full.shape[0] == 80 * (T - 1)
        np.testing.assert_allclose(full.numpy(), whole, atol=1e-5 * np.abs(whole).max() + 1e-9)
    else:
        assert full is None


def test_chunked_griffin_lim_halo_exchange():
    _spawn(_run_chunked_gl, 2)
