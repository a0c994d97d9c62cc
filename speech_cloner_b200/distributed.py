"""Multi-GPU layer of the DSP hot path: one process per GPU, torch.distributed.

The reference is single-process (its readers loop ``for i_sample in range(n_samples)``,
TIMIT_reader.py:169); these are the B200-native equivalents the north star prescribes:

* **Utterance sharding** for the front-end and for batched Griffin-Lim: utterances are independent
  (every reduction of ``calc_MFCC_input`` is per utterance, audio_lib.py:126, :157, :172, :221, :231,
  :235), so ranks need no data-path collective; only an optional final gather of the ragged
  feature buffers.  Shards are balanced by frame count (longest-processing-time greedy).
* **Time-chunked long-form Griffin-Lim** (the chapter-length call of test.py:148-168): rank r owns
  a contiguous range of output samples.  One iteration couples a sample only to audio within
  ``halo`` samples (``n_fft`` of frame reach plus one hop for the frame-pair partner), so the waveform
  state needs a halo exchange with the two neighbours (NCCL send / recv over NVLink).  The exchange is
  COMMUNICATION-AVOIDING: ranks swap a ``k * halo`` wide strip once every ``k`` iterations and recompute
  the shrinking overlap redundantly in between (``sc_griffinlim_chunk_run`` queues the ``k`` launches
  without touching the host), instead of 199 latency-bound exchanges.  Tiles, the de-emphasis scan and
  the ``mean|y|`` / ``realse`` sums all sit on whole-signal grids, so the assembled result is
  bit-identical to the single-GPU one for any rank count.
* The prologue (``realse`` power law: two scalar sums, audio_lib.py:293-296) and the epilogue
  (de-emphasis IIR carry across chunk boundaries and the ``mean|y|`` renormalisation, :301-306) are
  distributed too: one small all_gather / one neighbour message each.

The compute callables are injectable so that the host logic runs under ``gloo`` on CPU in the
tests (world_size 2) with oracle-based step functions.
"""
from __future__ import annotations

import ctypes as C
import math
import os

from typing import Callable, List, Optional, Sequence

import numpy as np

IIR_CHUNK = 256       # samples per chunk of the de-emphasis scan (csrc/gl_kernels.cuh kIirChunk)


# --------------------------------------------------------------------------- sharding
def bind_to_gpu_cpus(device_index: int) -> Optional[List[int]]:
    """Pin this process to the CPU cores NVML reports as local to GPU ``device_index`` (same NUMA node / PCIe root).

    One process per GPU moves its features through pinned host buffers; Linux places those pages on the NUMA node of
    the allocating thread, so a rank that runs on the far socket sends every download across the inter-socket link and
    the host-buffer (`e2e`) rate of an 8-GPU node collapses.  Call this before allocating pinned memory
    (``FrontendPipeline``).  Returns the cores now allowed, or ``None`` when NVML has nothing to say (single node,
    restricted container): the affinity is then left alone.  ``bind_report`` says which of the two happened and why.
    """
    return bind_report(device_index)["cores"]


def bind_report(device_index: int) -> dict:
    """``bind_to_gpu_cpus`` with the reason spelled out (bench.py prints it as ``config.cpu_binding_rank0``)."""
    rep = {"cores": None, "action": "none", "why": ""}
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        local = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1}
        have = set(os.sched_getaffinity(0))
        allowed = sorted(local & have)
        if not allowed:
            rep["why"] = "NVML reports no CPU local to the GPU inside this process's affinity mask"
        elif len(allowed) == len(have):
            rep["why"] = f"all {len(have)} host cores are local to the GPU (single NUMA node): nothing to bind"
        else:
            os.sched_setaffinity(0, allowed)
            rep.update(cores=allowed, action="bound", why=f"{len(allowed)} of {len(have)} cores are local to GPU {device_index}")
    except Exception as e:                                  # NVML missing / restricted container
        rep["why"] = f"NVML affinity unavailable ({type(e).__name__})"
    return rep


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
    allb = torch.empty(world * max(sizes), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(allb, flat, group=group)             # the one collective of the front-end path
    out = [None] * len(wavs)
    for r, s in enumerate(shards):
        o = r * max(sizes)
        for i in s:
            trip = []
            for w in widths:
                n = frames[i] * w
                trip.append(allb[o:o + n].reshape(frames[i], w))
                o += n
            out[i] = tuple(trip)
    return out


def gather_packed(local, sizes: Sequence[int], group=None):
    """all_gather of per-rank packed 1-D buffers of known sizes (``sizes[r]`` elements on rank r): one padded
    ``all_gather_into_tensor``; returns the list of per-rank views.  The "one final gather" of SURVEY.md §8(e)."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return [local[: sizes[0]]]
    m = max(int(s) for s in sizes)
    if local.shape[0] == m:
        pad = local
    else:
        pad = torch.zeros(m, dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    allb = torch.empty(world * m, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(allb, pad, group=group)
    return [allb[r * m: r * m + int(sizes[r])] for r in range(world)]


# ------------------------------------------------------------------- chunked Griffin-Lim
def chunk_geometry(n_fft: int = 400, hop_length: int = 80, plan=None):
    """(align_frames, halo_samples, halo_frames, sum_block_samples) of a time-chunked run (``sc_chunk_geometry``).

    Without a plan (CPU tests) the same formulas are evaluated here; ``tests/test_gpu_distributed.py`` checks that the
    two agree.
    """
    if plan is not None:
        vals = [C.c_int64(0) for _ in range(4)]
        from . import _lib
        _lib.check(plan._lib.sc_chunk_geometry(plan._h, *[C.byref(v) for v in vals]), "sc_chunk_geometry")
        return tuple(int(v.value) for v in vals)
    fast = (n_fft, hop_length) == (400, 80)
    tile_hops = 28 if fast else generic_tile_hops(n_fft, hop_length)
    align = tile_hops * (IIR_CHUNK // math.gcd(IIR_CHUNK, tile_hops * hop_length))
    halo_frames = -(-n_fft // hop_length) + 1
    return align, halo_frames * hop_length, halo_frames, align * hop_length


def generic_tile_hops(n_fft: int, hop: int) -> int:
    """Complete hops per tile of the generic-size iteration kernel (csrc/generic_kernels.cuh gen_gl_out_per_tile)."""
    return max(-(-n_fft // hop), 8)


def chunk_bounds(n_frames: int, world_size: int, hop_length: int = 80, align_frames: int = 112):
    """Split hop*(T-1) output samples into contiguous per-rank ranges.

    Cuts are multiples of ``align_frames`` hops (``sc_chunk_geometry``: the iteration kernel's tile grid and the
    256-sample grid of the de-emphasis scan) so that every rank works on whole-signal grids.  Returns a list of
    (lo, hi) sample ranges, one per rank (trailing ranks may be empty for short signals).
    """
    total = hop_length * (n_frames - 1)
    hops = n_frames - 1
    per = -(-hops // world_size)
    per = -(-per // align_frames) * align_frames
    cuts = [min(total, r * per * hop_length) for r in range(world_size)] + [total]
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


class ChunkedGriffinLim:
    """Time-chunked ``from_power_to_wav`` / Griffin-Lim of ONE long spectrogram across the ranks of a process group.

    Every rank passes its own rows of the time-major magnitude / initial phase (frames ``frame_range()``,
    halo included) and gets its chunk of the waveform back.  ``steps_per_exchange`` iterations run between two
    halo exchanges.  ``step`` injects a per-iteration compute callable with the signature of
    ``sc_griffinlim_chunk_step`` (CPU tests); by default ``sc_griffinlim_chunk_run`` queues a whole round.
    """

    def __init__(self, n_frames: int, hop_length: int = 80, n_fft: int = 400, group=None, step: Optional[Callable] = None,
                 steps_per_exchange: int = 20, plan=None, world: Optional[int] = None, rank: Optional[int] = None):
        dist = _dist()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world is not None:                 # planning / single-process emulation of a rank (tests drive the exchanges)
            self.world, self.rank = int(world), int(rank or 0)
        self.T, self.hop, self.n_fft = int(n_frames), int(hop_length), int(n_fft)
        self.total = self.hop * (self.T - 1)
        self.step = step
        self.plan = plan
        if plan is None and step is None:
            from .audio_lib import DspPlan
            self.plan = DspPlan.get(n_fft=self.n_fft, win_length=self.n_fft, hop_length=self.hop)
        self.align, self.halo, self.halo_frames, self.sum_block = chunk_geometry(self.n_fft, self.hop, self.plan)
        self.bounds = chunk_bounds(self.T, self.world, self.hop, self.align)
        self.lo, self.hi = self.bounds[self.rank]
        # a strip of k * halo samples must come from the direct neighbour: cap k by the shortest non-empty chunk
        shortest = min([b - a for a, b in self.bounds if b > a] or [self.halo])
        if self.world > 1 and shortest < self.halo:
            raise ValueError(f"chunk shorter than the {self.halo}-sample halo: use fewer ranks for this signal")
        self.k = max(1, min(int(steps_per_exchange), shortest // self.halo)) if self.world > 1 else 1 << 30

    def chunk_geometry_matches_host(self) -> bool:
        """The library's ``sc_chunk_geometry`` against the host formulas the CPU tests use."""
        return chunk_geometry(self.n_fft, self.hop, None) == (self.align, self.halo, self.halo_frames, self.sum_block)

    # -- ranges ---------------------------------------------------------------------------------------------------
    def _k_for(self, n_iters: int) -> int:
        return max(1, min(self.k, int(n_iters)))

    def frame_range(self, rank: Optional[int] = None, n_iters: int = 1 << 30):
        """Frames whose magnitude rows rank ``rank`` must hold: its hops plus k * halo_frames either side."""
        lo, hi = self.bounds[self.rank if rank is None else rank]
        k = self._k_for(n_iters) if self.world > 1 else 0
        return max(0, lo // self.hop - k * self.halo_frames), min(self.T, -(-hi // self.hop) + k * self.halo_frames + 1)

    def own_frame_range(self, rank: Optional[int] = None):
        """Frames this rank accounts for in whole-signal sums (``realse`` prologue): cut points / hop; the last
        non-empty rank also owns frame T-1."""
        r = self.rank if rank is None else rank
        lo, hi = self.bounds[r]
        last = max(i for i, (a, b) in enumerate(self.bounds) if b > a)
        if hi <= lo:
            return 0, 0
        return lo // self.hop, (self.T if r == last else hi // self.hop)

    def ext_range(self, rank: Optional[int] = None, n_iters: int = 1 << 30):
        lo, hi = self.bounds[self.rank if rank is None else rank]
        k = self._k_for(n_iters) if self.world > 1 else 0
        return max(0, lo - k * self.halo), min(self.total, hi + k * self.halo)

    def _peer(self, r):
        dist = _dist()
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def _live(self, r):
        return 0 <= r < self.world and self.bounds[r][1] > self.bounds[r][0]

    # -- halo exchange: my first / last `width` samples go to the left / right neighbour, theirs land in my halo
    def _exchange(self, buf, e_lo, width):
        if self.world == 1 or self.hi <= self.lo or width <= 0:
            return
        dist = _dist()
        ops = []
        lo, hi = self.lo, self.hi
        left, right = self.rank - 1, self.rank + 1
        if self._live(left):
            ops += [dist.P2POp(dist.isend, buf[lo - e_lo: lo - e_lo + width], self._peer(left), self.group),
                    dist.P2POp(dist.irecv, buf[lo - e_lo - width: lo - e_lo], self._peer(left), self.group)]
        if self._live(right):
            w_r = min(width, self.total - hi)
            ops += [dist.P2POp(dist.isend, buf[hi - e_lo - width: hi - e_lo], self._peer(right), self.group),
                    dist.P2POp(dist.irecv, buf[hi - e_lo: hi - e_lo + w_r], self._peer(right), self.group)]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    # -- Griffin-Lim ----------------------------------------------------------------------------------------------
    def run(self, amp_local, phase0_local, n_iters: int):
        """amp_local / phase0_local: rows ``frame_range(n_iters=n_iters)`` (time-major, float32).  Returns the
        float32 chunk ``[lo, hi)`` of the waveform after ``n_iters`` iterations (audio_lib.py:259)."""
        import torch
        n_iters = int(n_iters)
        if n_iters < 1:
            raise ValueError("n_iters must be >= 1")
        k = self._k_for(n_iters)
        f_lo, f_hi = self.frame_range(n_iters=n_iters)
        e_lo, e_hi = self.ext_range(n_iters=n_iters)
        if amp_local.shape[0] != f_hi - f_lo:
            raise ValueError("amp_local must hold the frame_range(n_iters=...) rows")
        dev = amp_local.device
        a = torch.zeros(e_hi - e_lo, dtype=torch.float32, device=dev)
        b = torch.zeros_like(a)
        if self.hi <= self.lo:
            return a[:0]
        done = 0
        while done < n_iters:
            n = min(k, n_iters - done)
            self._round(amp_local, phase0_local if done == 0 else None, f_lo, f_hi, a, b, e_lo, e_hi, n)
            if n & 1:
                a, b = b, a                                    # the state is always in `a` between rounds
            done += n
            if done < n_iters:
                self._exchange(a, e_lo, min(k, n_iters - done) * self.halo)
        return a[self.lo - e_lo: self.hi - e_lo]

    def _round(self, amp, phase0, f_lo, f_hi, a, b, e_lo, e_hi, n):
        """n iterations reading `a` first; window of step m = chunk widened by (n-1-m) halos."""
        if self.step is None:
            import torch
            from . import _lib
            rc = self.plan._lib.sc_griffinlim_chunk_run(
                self.plan._h, amp.data_ptr(), phase0.data_ptr() if phase0 is not None else None, f_lo, f_hi - f_lo, self.T,
                a.data_ptr(), b.data_ptr(), e_lo, e_hi - e_lo, self.lo, self.hi - self.lo, n, self.halo,
                torch.cuda.current_stream().cuda_stream)
            _lib.check(rc, "sc_griffinlim_chunk_run")
            return
        src, dst = a, b
        for m in range(n):
            grow = (n - 1 - m) * self.halo
            lo, hi = max(e_lo, self.lo - grow), min(e_hi, self.hi + grow)
            first = phase0 is not None and m == 0
            self.step(amp, phase0 if first else None, f_lo, f_hi - f_lo, self.T, None if first else src, e_lo, e_hi - e_lo,
                      dst[lo - e_lo: hi - e_lo], lo, hi - lo)
            src, dst = dst, src

    # -- prologue (audio_lib.py:290-298) --------------------------------------------------------------------------
    def power_to_amp(self, p_local, P_dB_norm_factor: float = 0.01, realse: float = 1.0, n_iters: int = 1 << 30):
        """Rows ``frame_range()`` of the normalised power-dB map -> magnitudes, in a new buffer.  With ``realse != 1``
        the two means span the whole signal: block partials of the rows each rank owns are all_gathered and summed
        in block order by every rank (bit-identical to one GPU)."""
        import torch
        from . import _lib
        lib, h = self.plan._lib, self.plan._h
        st = torch.cuda.current_stream().cuda_stream
        f_lo, f_hi = self.frame_range(n_iters=n_iters)
        if p_local.shape[0] != f_hi - f_lo:
            raise ValueError("p_local must hold the frame_range(n_iters=...) rows")
        amp = torch.empty_like(p_local)
        allp, n_blocks = None, 0
        if realse != 1.0:
            counts = []
            for r in range(self.world):
                o_lo, o_hi = self.own_frame_range(r)
                counts.append(2 * (-(-(o_hi - o_lo) // self.align)))
            o_lo, o_hi = self.own_frame_range()
            part = torch.zeros(counts[self.rank], dtype=torch.float64, device=p_local.device)
            if o_hi > o_lo:
                _lib.check(lib.sc_p2a_chunk_partial(h, p_local[o_lo - f_lo:].data_ptr(), o_hi - o_lo, float(realse),
                                                    part.data_ptr(), st), "sc_p2a_chunk_partial")
            allp = torch.cat(gather_packed(part, counts, self.group)).contiguous()
            n_blocks = allp.shape[0] // 2
        if f_hi > f_lo:
            _lib.check(lib.sc_p2a_chunk_apply(h, p_local.data_ptr(), f_hi - f_lo, float(P_dB_norm_factor), float(realse),
                                              allp.data_ptr() if allp is not None else None, n_blocks, amp.data_ptr(), st),
                       "sc_p2a_chunk_apply")
        return amp

    # -- epilogue (audio_lib.py:301-306) --------------------------------------------------------------------------
    def deemph_renorm(self, chunk, pre_emphasis: float = 0.97, mean_abs_amp_norm: float = 0.01):
        """float32 chunk ``[lo, hi)`` -> float64 chunk of ``y * (m / mean|y|)`` after the de-emphasis IIR.

        Per boundary ONE message (the left neighbour's last ``win`` chunk responses of the scan), then ONE all_gather of
        the per-block sums of |y|; both are summed in a fixed whole-signal order, so any rank count gives the same bits.
        """
        import torch
        from . import _lib
        dist = _dist()
        lib, h = self.plan._lib, self.plan._h
        st = torch.cuda.current_stream().cuda_stream
        dev = chunk.device
        n = self.hi - self.lo
        win = int(lib.sc_deemph_chunk_window(float(pre_emphasis)))
        if win == 0 and self.world > 1:
            raise NotImplementedError("pre_emphasis is too close to 1 for the windowed carry: de-emphasise on one GPU")
        n_chunks = -(-n // IIR_CHUNK)
        loc = torch.zeros(win + n_chunks, dtype=torch.float64, device=dev)
        if self.world == 1 and win == 0:
            from .audio_lib import _GlLayout
            out = torch.empty(n, dtype=torch.float64, device=dev)
            _lib.check(lib.sc_deemph_renorm_batch(h, chunk.data_ptr(), _lib.i64_array([0, n]), _lib.i64_array([n]), 1,
                                                  float(pre_emphasis), float(mean_abs_amp_norm), out.data_ptr(), st), "deemph")
            return out
        if n > 0:
            _lib.check(lib.sc_deemph_chunk_local(h, chunk.data_ptr(), self.lo, n, self.total, float(pre_emphasis),
                                                 loc[win:].data_ptr(), st), "sc_deemph_chunk_local")
        if self.world > 1 and n > 0:
            ops = []
            left, right = self.rank - 1, self.rank + 1
            if self._live(right):
                if n_chunks < win:
                    raise ValueError("chunk shorter than the de-emphasis carry window: use fewer ranks")
                ops.append(dist.P2POp(dist.isend, loc[win + n_chunks - win:], self._peer(right), self.group))
            if self._live(left):
                ops.append(dist.P2POp(dist.irecv, loc[:win], self._peer(left), self.group))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
        out = torch.empty(n, dtype=torch.float64, device=dev)
        counts = [-(-(b - a) // self.sum_block) for a, b in self.bounds]
        sums = torch.zeros(counts[self.rank], dtype=torch.float64, device=dev)
        if n > 0:
            _lib.check(lib.sc_deemph_chunk_apply(h, chunk.data_ptr(), self.lo, n, self.total, float(pre_emphasis),
                                                 loc.data_ptr(), win, out.data_ptr(), sums.data_ptr(), st), "sc_deemph_chunk_apply")
        alls = torch.cat(gather_packed(sums, counts, self.group)).contiguous()
        if n > 0:
            _lib.check(lib.sc_renorm_chunk(h, out.data_ptr(), n, alls.data_ptr(), alls.shape[0], self.total,
                                           float(mean_abs_amp_norm), st), "sc_renorm_chunk")
        return out

    def from_power_to_wav(self, p_local, phase0_local, P_dB_norm_factor=0.01, pre_emphasis=0.97, mean_abs_amp_norm=0.01,
                          n_iter=200, realse=1.0):
        """The whole ``from_power_to_wav`` (audio_lib.py:278-308) on this rank's chunk; returns float64 ``[lo, hi)``."""
        amp = self.power_to_amp(p_local, P_dB_norm_factor, realse, n_iters=n_iter)
        chunk = self.run(amp, phase0_local, n_iter)
        return self.deemph_renorm(chunk, pre_emphasis, mean_abs_amp_norm)

    def gather(self, chunk, dst: Optional[int] = 0):
        """Final gather of the per-rank chunks into the whole waveform on rank ``dst`` (None elsewhere; ``dst=None``:
        on every rank)."""
        import torch
        if self.world == 1:
            return chunk
        sizes = [b - a for a, b in self.bounds]
        parts = gather_packed(chunk, sizes, self.group)
        if not (dst is None or self.rank == dst):
            return None
        return torch.cat(parts)
