"""2-GPU NCCL tests of the multi-GPU paths (skipped with fewer than 2 devices; run with gpurun --gpus 2).

Chunked == unchunked Griffin-Lim is the distributed correctness test of SURVEY.md §4 / §8(e).
"""
import os
import socket

import numpy as np
import pytest
import torch

from speech_cloner_b200 import synth

pytestmark = pytest.mark.gpu
HP = dict(synth.HP_ENC)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank, world, port, fn):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        fn(rank, world)
    finally:
        dist.destroy_process_group()


def _need2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def _featurize(rank, world):
    from speech_cloner_b200 import audio_lib as al
    from speech_cloner_b200 import distributed as D
    wavs = [synth.utterance(900 + i, s) for i, s in enumerate([1.0, 0.4, 2.0, 0.7, 1.3, 0.2, 0.9])]
    got = D.featurize_sharded(wavs, gather=True, return_device=True, **HP)
    want = al.calc_MFCC_input_batch(wavs, return_device=True, **HP)
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            assert torch.equal(a, b)


def _chunked(rank, world):
    """Chunked == unchunked, bit for bit, through the public distributed path (NCCL halo exchange every k iterations,
    distributed prologue with realse != 1, windowed IIR carry across the boundary, gathered |y| sums)."""
    from oracle import audio_lib_oracle as oracle
    from speech_cloner_b200 import audio_lib as al
    from speech_cloner_b200 import distributed as D
    T = 2001
    base = oracle.calc_MFCC_input(synth.utterance(31, 4.0), **HP)[2]
    P = np.concatenate([base] * 8)[:T]
    np.random.seed(4)
    ph = np.pi * np.random.rand(201, T)
    n_iter = 25
    for realse, k in ((1.0, 20), (1.2, 4), (1.0, 1)):
        whole = al.from_power_to_wav(P, P_dB_norm_factor=0.01, pre_emphasis=0.97, hop_length=80, win_length=400,
                                     mean_abs_amp_norm=0.045, n_iter=n_iter, realse=realse, verbose=False, phase0=ph)
        gl = D.ChunkedGriffinLim(T, 80, 400, steps_per_exchange=k)
        assert gl.chunk_geometry_matches_host()
        f_lo, f_hi = gl.frame_range(n_iters=n_iter)
        p_loc = torch.from_numpy(np.ascontiguousarray(P[f_lo:f_hi])).cuda()
        ph_loc = torch.from_numpy(np.ascontiguousarray(ph.T[f_lo:f_hi]).astype(np.float32)).cuda()
        y = gl.from_power_to_wav(p_loc, ph_loc, P_dB_norm_factor=0.01, pre_emphasis=0.97, mean_abs_amp_norm=0.045,
                                 n_iter=n_iter, realse=realse)
        assert y.dtype == torch.float64
        np.testing.assert_array_equal(y.cpu().numpy(), whole[gl.lo:gl.hi])   # bit-identical to the single-GPU run
        full = gl.gather(y, dst=0)
        if rank == 0:
            np.testing.assert_array_equal(full.cpu().numpy(), whole)
        else:
            assert full is None


def test_featurize_sharded_nccl():
    _need2()
    torch.multiprocessing.spawn(_entry, args=(2, _free_port(), _featurize), nprocs=2, join=True)


def test_chunked_griffin_lim_nccl_bit_identical():
    _need2()
    torch.multiprocessing.spawn(_entry, args=(2, _free_port(), _chunked), nprocs=2, join=True)


def test_plan_belongs_to_its_device():
    """a plan's tables, function attributes and SM count belong to the device it was created on; using it with
    another device current is refused (SC_ERR_INVALID -> ValueError), and the public API builds one plan per device."""
    _need2()
    from speech_cloner_b200 import audio_lib as al
    y = synth.utterance(11, 1.0)
    with torch.cuda.device(0):
        want = al.calc_MFCC_input(y, **HP)
        plan0 = al.DspPlan(sr=16000, n_fft=400, win_length=400, hop_length=80, n_mels=80, n_mfcc=40)
    with torch.cuda.device(1):
        got = al.calc_MFCC_input(y, **HP)                      # second device: its own plan, same results
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a, b)
        lay = al.FrontendLayout([len(y)], 80)
        dev = torch.from_numpy(y).cuda()
        with pytest.raises(ValueError):
            al.frontend_device(plan0, dev, lay)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
