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
    from oracle import audio_lib_oracle as oracle
    from speech_cloner_b200 import distributed as D
    from speech_cloner_b200.audio_lib import DspPlan, _GlLayout, griffin_lim_device
    T = 2001
    base = oracle.calc_MFCC_input(synth.utterance(31, 4.0), **HP)[2]
    P = np.concatenate([base] * 8)[:T]
    amp = torch.from_numpy(np.sqrt(np.power(np.float32(10.0), np.float32(0.1) * (P / np.float32(0.01) - np.float32(80.0))))
                           .astype(np.float32)).cuda()
    np.random.seed(4)
    ph = torch.from_numpy((np.pi * np.random.rand(T, 201)).astype(np.float32)).cuda()
    n_iter = 25
    plan = DspPlan.get(n_fft=400, win_length=400, hop_length=80)
    whole = griffin_lim_device(plan, amp, ph, _GlLayout([T], 80), n_iter)[: 80 * (T - 1)].clone()
    gl = D.ChunkedGriffinLim(T, 80, 400)
    f_lo, f_hi = gl.frame_range()
    chunk = gl.run(amp[f_lo:f_hi].contiguous(), ph[f_lo:f_hi].contiguous(), n_iter)
    assert torch.equal(chunk, whole[gl.lo:gl.hi])                 # bit-identical to the single-GPU run
    full = gl.gather(chunk, dst=0)
    if rank == 0:
        assert torch.equal(full, whole)


def test_featurize_sharded_nccl():
    _need2()
    torch.multiprocessing.spawn(_entry, args=(2, _free_port(), _featurize), nprocs=2, join=True)


def test_chunked_griffin_lim_nccl_bit_identical():
    _need2()
    torch.multiprocessing.spawn(_entry, args=(2, _free_port(), _chunked), nprocs=2, join=True)
