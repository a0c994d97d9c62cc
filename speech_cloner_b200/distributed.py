"""Multi-GPU layer of the DSP hot path: one process per GPU, torch.distributed.

The reference is single-process (its readers loop ``for i_sample in range(n_samples)``,
TIMIT_reader.py:169); these are the B200-native equivalents the north star prescribes:

* **Utterance sharding** for the front-end and for batched Griffin-Lim: utterances are independent
  (every reduction of ``calc_MFCC_input`` is per utterance, audio_lib.py:126, :157, :172, :221, :231,
  :235), so ranks need no data-path collective; only an optional final gather of the ragged
  feature buffers.  Shards are balanced by frame count (longest-processing-time greedy).
* **Time-chunked long-form Griffin-Lim** (the chapter-length call of test.py:148-168): rank r owns
  a contiguous range of output samples; one iteration couples a sample only to audio within
  +-480 samples (frames t-2..t+3 of its hop, each 400 long), so per iteration the neighbours swap
  a 480-sample halo (1.9 KB) with ``batch_isend_irecv`` over NVLink.  Tiles sit on a whole-signal
  grid, which makes the chunked result bit-identical to the single-GPU one.

The compute callables are injectable so that the host logic runs under ``gloo`` on CPU in the
tests (world_size 2) with an oracle-based step function.
"""
from __future__ import annotations

import os

from typing import Callable, List, Optional, Sequence

import numpy as np

HALO = 480          # samples: 2.5 hops of frame reach + 200 of half window, rounded up to 6 hops
HALO_FRAMES = 6


# --------------------------------------------------------------------------- sharding
def bind_to_gpu_cpus(device_index: int) -> Optional[List[int]]:
    """Pin this process to the CPU cores NVML reports as local to GPU ``device_index`` (same NUMA node / PCIe root).

    One process per GPU moves its features through pinned host buffers; Linux places those pages on the NUMA node of
    the allocating thread, so a rank that runs on the far socket sends every download across the inter-socket link and
    the host-buffer (`e2e`) rate of an 8-GPU node collapses.  Call this before allocating pinned memory
    (``FrontendPipeline``).  Returns the cores now allowed, or ``None`` when NVML has nothing to say (single node,
    restricted container): the affinity is then left alone.
    """
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        local = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1}
        allowed = sorted(local & set(os.sched_getaffinity(0)))
        if not allowed or len(allowed) == len(os.sched_getaffinity(0)):
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def shard_by_frames(lengths: Sequence[int], world_size: int, hop_length: int = 80) -> List[List[int]]:
    """Longest-processing-time greedy partition of utterance indices, balanced by frame count.

    Deterministic (ties broken by index) so every rank computes the same partition locally.
    """
    frames = [1 + int(n) // hop_length for n in lengths]
    order = sorted(range(len(lengths)), key=lambda i: (-frames[i], i))
    load = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += frames[i]
    for s in shards:
        s.sort()
    return shards


def _dist():
    import torch.distributed as dist
    return dist


def featurize_sharded(wavs, compute: Optional[Callable] = None, gather: bool = True, group=None, **hp):
    """Front-end over utterances sharded across the ranks of ``group``.

    ``compute(list_of_wavs, **hp) -> list of (MFCC, M_dB, P_dB)`` defaults to the CUDA batch call.
    Returns, on every rank when ``gather`` (all_gather_object-free: sizes are implied by the
    lengths every rank already knows), the full list in the original utterance order; otherwise
    a dict {utterance index: triple} of the local shard.
    """
    import torch
    dist = _dist()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if compute is None:
        from . import audio_lib
        compute = audio_lib.calc_MFCC_input_batch
    hop = int(hp.get("hop_length", 40))
    shards = shard_by_frames([len(w) for w in wavs], world, hop)
    mine = shards[rank]
    local = compute([wavs[i] for i in mine], **hp) if mine else []
    if not gather or world == 1:
        res = {i: t for i, t in zip(mine, local)}
        return [res[i] for i in range(len(wavs))] if world == 1 else res

    # ---- final gather: one flat float32 buffer per rank, sizes known from the lengths
    frames = [1 + len(w) // hop for w in wavs]
    if local:
        widths = [int(a.shape[1]) for a in local[0]]
    else:
        widths = [0, 0, 0]
    wt = torch.tensor(widths, dtype=torch.int64)
    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    wt = wt.to(dev)
    dist.all_reduce(wt, op=dist.ReduceOp.MAX, group=group)          # ranks with empty shards learn the widths
    widths = [int(x) for x in wt.tolist()]
    row = sum(widths)
    sizes = [sum(frames[i] for i in s) * row for s in shards]
    flat = torch.zeros(max(sizes), dtype=torch.float32, device=dev)      # padded to the largest shard
    o = 0
    for t in local:
        for a in t:
            a = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
            flat[o:o + a.numel()] = a.reshape(-1).to(dev)
            o += a.numel()
    bufs = [torch.empty(max(sizes), dtype=torch.float32, device=dev) for _ in sizes]
    dist.all_gather(bufs, flat, group=group)                         # the one collective of the front-end path
    out = [None] * len(wavs)
    for r, s in enumerate(shards):
        o = 0
        b = bufs[r]
        for i in s:
            trip = []
            for w in widths:
                n = frames[i] * w
                trip.append(b[o:o + n].reshape(frames[i], w))
                o += n
            out[i] = tuple(trip)
    return out


# ------------------------------------------------------------------- chunked Griffin-Lim
def chunk_bounds(n_frames: int, world_size: int, hop_length: int = 80, align_frames: int = 28):
    """Split hop*(T-1) output samples into contiguous per-rank ranges.

    Cuts are multiples of ``align_frames`` hops (the kernel's 28-hop tile) so no tile straddles
    two ranks more than necessary.  Returns a list of (lo, hi) sample ranges, one per rank.
    """
    total = hop_length * (n_frames - 1)
    hops = n_frames - 1
    per = -(-hops // world_size)
    per = -(-per // align_frames) * align_frames
    cuts = [min(total, r * per * hop_length) for r in range(world_size)] + [total]
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


class ChunkedGriffinLim:
    """Time-chunked Griffin-Lim of ONE long spectrogram across the ranks of a process group.

    Every rank passes its own rows of the time-major magnitude / initial phase (frames
    ``frame_range(rank)``, halo included) and gets its chunk of the waveform back.
    ``step`` is the per-iteration compute callable with the signature of
    ``sc_griffinlim_chunk_step`` (defaults to the CUDA one).
    """

    def __init__(self, n_frames: int, hop_length: int = 80, n_fft: int = 400, group=None, step: Optional[Callable] = None):
        dist = _dist()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.T, self.hop, self.n_fft = int(n_frames), int(hop_length), int(n_fft)
        self.total = self.hop * (self.T - 1)
        self.bounds = chunk_bounds(self.T, self.world, self.hop)
        self.lo, self.hi = self.bounds[self.rank]
        self.step = step

    def frame_range(self, rank: Optional[int] = None):
        """Frames whose magnitude rows rank ``rank`` must hold: its hops plus HALO_FRAMES either side."""
        lo, hi = self.bounds[self.rank if rank is None else rank]
        return max(0, lo // self.hop - HALO_FRAMES), min(self.T, -(-hi // self.hop) + HALO_FRAMES + 1)

    def ext_range(self, rank: Optional[int] = None):
        lo, hi = self.bounds[self.rank if rank is None else rank]
        return max(0, lo - HALO), min(self.total, hi + HALO)

    # -- halo exchange: my first / last HALO samples go to the left / right neighbour
    def _exchange(self, ext, e_lo):
        if self.world == 1:
            return
        dist = _dist()
        import torch
        ops, keep = [], []
        lo, hi = self.lo, self.hi
        left, right = self.rank - 1, self.rank + 1
        empty = hi <= lo
        if left >= 0 and not empty and self.bounds[left][1] > self.bounds[left][0]:
            n_send = min(HALO, hi - lo)
            send = ext[lo - e_lo: lo - e_lo + n_send].contiguous()
            n_recv = lo - e_lo
            recv = torch.empty(n_recv, dtype=ext.dtype, device=ext.device)
            ops += [dist.P2POp(dist.isend, send, self._peer(left), self.group),
                    dist.P2POp(dist.irecv, recv, self._peer(left), self.group)]
            keep.append((recv, 0, n_recv))
        if right < self.world and not empty and self.bounds[right][1] > self.bounds[right][0]:
            n_send = min(HALO, hi - lo)
            send = ext[hi - e_lo - n_send: hi - e_lo].contiguous()
            n_recv = min(self.total, hi + HALO) - hi
            recv = torch.empty(n_recv, dtype=ext.dtype, device=ext.device)
            ops += [dist.P2POp(dist.isend, send, self._peer(right), self.group),
                    dist.P2POp(dist.irecv, recv, self._peer(right), self.group)]
            keep.append((recv, hi - e_lo, n_recv))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for recv, off, n in keep:
            ext[off:off + n] = recv[:n]

    def _peer(self, r):
        dist = _dist()
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def run(self, amp_local, phase0_local, n_iters: int):
        """amp_local / phase0_local: rows ``frame_range()`` (time-major, float32).  Returns the
        float32 chunk ``[lo, hi)`` of the waveform after ``n_iters`` iterations (audio_lib.py:259)."""
        import torch
        f_lo, f_hi = self.frame_range()
        e_lo, e_hi = self.ext_range()
        assert amp_local.shape[0] == f_hi - f_lo, "amp_local must hold frame_range() rows"
        dev = amp_local.device
        cur = torch.zeros(e_hi - e_lo, dtype=torch.float32, device=dev)
        nxt = torch.zeros_like(cur)
        # chunks shorter than the halo would need multi-hop exchanges; refuse instead of being wrong
        for (a, b) in self.bounds:
            if 0 < b - a < HALO and self.world > 1:
                raise ValueError("chunk shorter than the 480-sample halo: use fewer ranks for this signal")
        step = self.step or _cuda_step(self.n_fft, self.hop)
        for it in range(int(n_iters)):
            step(amp_local, phase0_local if it == 0 else None, f_lo, f_hi - f_lo, self.T,
                 cur if it else None, e_lo, e_hi - e_lo, nxt[self.lo - e_lo: self.hi - e_lo], self.lo, self.hi - self.lo)
            if it != n_iters - 1:
                self._exchange(nxt, e_lo)
            cur, nxt = nxt, cur
        return cur[self.lo - e_lo: self.hi - e_lo]

    def gather(self, chunk, dst: int = 0):
        """Final gather of the per-rank chunks into the whole waveform on rank ``dst`` (None elsewhere)."""
        import torch
        if self.world == 1:
            return chunk
        dist = _dist()
        sizes = [b - a for a, b in self.bounds]
        pad = torch.zeros(max(sizes), dtype=chunk.dtype, device=chunk.device)
        pad[: chunk.shape[0]] = chunk
        bufs = [torch.empty_like(pad) for _ in sizes]
        dist.all_gather(bufs, pad, group=self.group)
        if not (self.rank == dst or dst is None):
            return None
        return torch.cat([b[:n] for b, n in zip(bufs, sizes)])


def _cuda_step(n_fft: int, hop: int):
    """``sc_griffinlim_chunk_step`` bound to a plan (include/speechdsp.h)."""
    import torch
    from . import _lib
    from .audio_lib import DspPlan
    plan = DspPlan.get(n_fft=n_fft, win_length=n_fft, hop_length=hop)
    lib = _lib.load()

    def step(amp, phase0, first_frame, n_local, n_total, wav_in, wav_first, wav_count, wav_out, out_first, out_count):
        if out_count <= 0:
            return
        rc = lib.sc_griffinlim_chunk_step(
            plan._h, amp.data_ptr(), phase0.data_ptr() if phase0 is not None else None, first_frame, n_local, n_total,
            wav_in.data_ptr() if wav_in is not None else None, wav_first, wav_count, wav_out.data_ptr(), out_first,
            out_count, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "sc_griffinlim_chunk_step")
    return step
