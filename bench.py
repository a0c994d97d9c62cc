#!/usr/bin/env python
"""bench.py — audio-seconds/second of the speech-cloner DSP hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--precision fp64|fp32] [--no-gl] [--no-cpu]

A "step" is one pass of the front-end (calc_MFCC_input, /root/reference/audio_lib.py:89-244) over
BASELINE.json configs[1]: 256 synthetic ARCTIC-shaped utterances x 4 s at the hp/*.json settings
(205 056 frames, 65.5 MB in, 296.1 MB out).  One JSON line is printed by rank 0:

  value      whole-job audio-s/s with the waveforms already resident in HBM (CUDA events, max over ranks)
  e2e        the same through the host-buffer pipeline: pinned H2D of every waveform and pinned D2H of the
             three feature arrays inside the timed region
  roofline   achieved algorithmic GB/s (1 764 B/frame) over the summed kernel time of a
             step, against MEASURED_PEAKS.json hbm_gbs; per-kernel shares from CUDA events in the library
  cpu_baseline   the CPU oracle (restated reference path) on this box's host cores, bounded sample
  griffin_lim    secondary measurement, configs[2]: 64 spectrograms x 5 s, 200 iterations (test.py:87)

Under torchrun every rank featurises its own copy of the batch (weak scaling, no data-path
collective: utterances are independent).  `--impl reference` times the CPU oracle with
all host cores on the same config (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

# the CPU legs count cores explicitly: one BLAS/OpenMP thread per process (set before numpy loads)
for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_k, "1")
# stdout carries exactly one JSON line: everything else that libraries write to file descriptor 1 (NCCL's version
# banner, ...) is sent to stderr, and the line itself goes to the saved descriptor (emit()).
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj) -> None:
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_UTTS, SECONDS, SR = 256, 4.0, 16000
GL_UTTS, GL_SECONDS, GL_ITERS = 64, 5.0, 200
FE_BYTES_PER_FRAME = 4 * (80 + 201 + 80 + 80)            # wav in + three outputs
GL_BYTES_PRIMARY = 4 * 201 + 8 * 201 + 8 * 201           # complex64 spectrogram state
GL_BYTES_STRICT = 4 * 201 + 2 * 4 * 80                   # waveform state (what the kernel moves)


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, windows):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.samples:
            if not any(a <= t <= b for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------ CPU legs
def _oracle_utt(args):
    from oracle import audio_lib_oracle as oracle
    from speech_cloner_b200 import synth
    y = args
    return oracle.calc_MFCC_input(y, **synth.HP_ENC)[0].shape[0]


def cpu_frontend_single(wavs):
    """Sequential single-process loop exactly like TIMIT_reader.py:169."""
    from oracle import audio_lib_oracle as oracle
    from speech_cloner_b200 import synth
    t = time.perf_counter()
    for y in wavs:
        oracle.calc_MFCC_input(y, **synth.HP_ENC)
    dt = time.perf_counter() - t
    return sum(len(y) for y in wavs) / SR / dt


def cpu_gl_single(P, phase0, iters):
    from oracle import audio_lib_oracle as oracle
    t = time.perf_counter()
    oracle.from_power_to_wav(P, P_dB_norm_factor=0.01, pre_emphasis=0.97, hop_length=80, win_length=400,
                             mean_abs_amp_norm=0.045, n_iter=iters, realse=1.0, verbose=False, phase0=phase0)
    return time.perf_counter() - t


def run_reference(args):
    """`--impl reference`: the reference's CPU path (oracle restatement; librosa is not installable) with
    every host core, one utterance per task, BLAS threads pinned to 1."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from speech_cloner_b200 import synth
    cores = os.cpu_count() or 1
    sample_n = 64
    wavs = synth.batch(2, sample_n, SECONDS)
    audio_s = sample_n * SECONDS
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_oracle_utt, wavs, chunksize=1)
        t = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_oracle_utt, wavs, chunksize=1)
        dt = (time.perf_counter() - t) / args.steps
    val = audio_s / dt
    sample = f"{sample_n} of the 256 utterances x {SECONDS:g} s per step, multiprocessing.Pool({cores})"
    emit({
        "impl": "reference", "metric": "audio-seconds/second, front-end (calc_MFCC_input)", "value": val,
        "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "frontend configs[1]: 256 x 4 s ARCTIC-shaped, hp/ds_dec_cfg_d.json", "sample": sample},
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------ GPU legs
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--no-gl", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--gl-long", action="store_true", help="also time the chapter-length chunked Griffin-Lim (configs[3])")
    ap.add_argument("--sweep", action="store_true", help="also time the 10 h featurization sweep (configs[4]), utterance-sharded")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer pipeline leg (profiling runs)")
    ap.add_argument("--no-probe", action="store_true", help="skip the 1.2 s clock probe loop (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from speech_cloner_b200 import audio_lib as al
    from speech_cloner_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cpu_bind = None
    if world > 1:
        # one process per GPU: keep the rank (and the pinned buffers it is about to allocate) on the GPU's NUMA node
        from speech_cloner_b200 import distributed as D
        cpu_bind = D.bind_to_gpu_cpus(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, peak_src = hbm_peak()
    hp = dict(synth.HP_ENC)
    plan_kw = dict(sr=hp["sr"], n_fft=400, win_length=400, hop_length=80, n_mels=80, n_mfcc=40, window="hann",
                   pre_emphasis=0.97, mfcc_normaleze_first_mfcc=True, mfcc_norm_factor=0.01, calc_mfcc_derivate=True,
                   M_dB_norm_factor=0.01, P_dB_norm_factor=0.01, mean_abs_amp_norm=0.003, clip_output=True,
                   fft_precision=args.precision)

    # ---- synthetic batch (every rank its own seeds: weak scaling)
    wavs = synth.batch(2 + 10 * rank, N_UTTS, SECONDS)
    audio_s = N_UTTS * SECONDS
    pipe = al.FrontendPipeline([len(w) for w in wavs], n_chunks=8, n_streams=3, **plan_kw)
    pipe.load(wavs)
    lay = pipe.layout
    frames = sum(lay.frames)
    plan = al.DspPlan(**plan_kw)
    wav_dev = pipe.wav_host.to("cuda")
    out = tuple(torch.empty_like(o) for o in pipe.out_dev)

    sampler = ClockSampler(local) if rank == 0 else None
    windows = []

    # ---- device-resident value
    for _ in range(args.warmup):
        al.frontend_device(plan, wav_dev, lay, out)
    barrier()
    al.launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        al.frontend_device(plan, wav_dev, lay, out)
    e1.record()
    barrier()
    windows.append((w0, time.time()))
    launches = al.launch_count()
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    value = world * audio_s / (ms_step * 1e-3)

    # ---- per-kernel times (library events), separate from the timed region above
    plan.profile(True)
    prof = np.zeros(3)
    n_prof = min(args.steps, 50)
    for _ in range(n_prof):
        al.frontend_device(plan, wav_dev, lay, out)
        prof += np.array(plan.profile_read()[:3])
    plan.profile(False)
    prof /= n_prof
    kern_ms = float(prof.sum())
    names = ["gain: k_fe_setup + k_abs_pairwise4 + k_gain_finalize", "pass A: k_fe_pass_a_ws", "pass B: k_fe_c00 + k_fe_pass_b3"]
    top = int(np.argmax(prof))
    fe_bytes = FE_BYTES_PER_FRAME * frames
    try:                                   # DRAM bytes per launch from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            traffic_tab = json.load(f)
    except Exception:
        traffic_tab = {}
    achieved = fe_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": (sum(traffic_tab.get(n, float("nan")) for n in names) if traffic_tab else None),
        "traffic_note": "DRAM read+write bytes of one step (all kernels), ncu --set full, profiles/r01_traffic.json; "
                        "algorithmic bytes of the step: %d" % fe_bytes,
        "kernel_traffic": traffic_tab.get(names[top]),
        "peak_source": peak_src,
        "definition": "1764 B/frame x frames of one step / summed device time of the step's kernels",
        "kernel": names[top],
        # the dominant kernel against its own algorithmic bytes: gain reads the audio once
        # (320 B/frame), pass A reads it again and writes the raw dB tiles (320 + 1124), pass B re-reads and rewrites
        # (1124 + 1444)
        "kernel_algorithmic_bytes_per_frame": [320, 1444, 2568][top],
        "kernel_achieved": [320, 1444, 2568][top] * frames / (float(prof[top]) * 1e-3) / 1e9,
        "kernel_frac": [320, 1444, 2568][top] * frames / (float(prof[top]) * 1e-3) / 1e9 / peak,
        "kernels": {n: {"ms": float(m), "share": float(m / kern_ms)} for n, m in zip(names, prof)},
        "step_ms_back_to_back": ms_step,
    }

    # ---- opt-in float32-FFT mode of the same step (documented accuracy: < 1e-4 abs, >= 99.7 % inside 1e-4/1e-5)
    alt = None
    if args.precision == "fp64":
        kw32 = dict(plan_kw); kw32["fft_precision"] = "fp32"
        plan32 = al.DspPlan(**kw32)
        for _ in range(args.warmup):
            al.frontend_device(plan32, wav_dev, lay, out)
        barrier()
        a32, b32 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a32.record()
        for _ in range(args.steps):
            al.frontend_device(plan32, wav_dev, lay, out)
        b32.record()
        barrier()
        ms32 = max_over_ranks(a32.elapsed_time(b32) / args.steps)
        alt = {"fft_precision": "fp32", "ms_per_step": ms32, "value": world * audio_s / (ms32 * 1e-3), "unit": "audio-s/s",
               "roofline_frac": fe_bytes / (ms32 * 1e-3) / 1e9 / peak,
               "note": "not the default: bins 70-80 dB below the utterance maximum can deviate by up to ~5e-5"}
        del plan32

    # ---- clock probe: keep the same step running ~1.2 s so nvidia-smi (50 ms period) sees it under load
    if rank == 0 and not args.no_probe:
        w0 = time.time()
        while time.time() - w0 < 1.2:
            for _ in range(50):
                al.frontend_device(plan, wav_dev, lay, out)
            torch.cuda.synchronize()
        windows.append((w0, time.time()))

    # ---- e2e through the host-buffer pipeline (pinned H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        for _ in range(3):
            pipe.run()
        barrier()
        e2e_steps = max(3, min(args.steps, 20))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pipe.run()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        e2e = {"value": world * audio_s / e2e_s, "unit": "audio-s/s", "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": int(pipe.h2d_bytes), "d2h_bytes_per_step": int(pipe.d2h_bytes),
               "path": "FrontendPipeline.run(): 8 utterance chunks over 3 streams, pinned host buffers"}

    # ---- Griffin-Lim, configs[2]
    gl = None
    if not args.no_gl:
        gl_wavs = synth.batch(3 + 10 * rank, GL_UTTS, GL_SECONDS)
        feats = al.calc_MFCC_input_batch(gl_wavs, return_device=True, **hp)
        Ps = [f[2][:1000].contiguous() for f in feats]                     # (1000, 201) decoder-shaped
        glay = al._GlLayout([1000] * GL_UTTS, 80)
        gplan = al.DspPlan.get(n_fft=400, win_length=400, hop_length=80)
        p_dev = torch.zeros((glay.frame_offsets[-1], 201), dtype=torch.float32, device="cuda")
        ph_dev = torch.zeros_like(p_dev)
        for i, (P, o) in enumerate(zip(Ps, glay.frame_offsets)):
            p_dev[o:o + 1000] = P
            np.random.seed(3000 + i)
            ph_dev[o:o + 1000] = torch.from_numpy((np.pi * np.random.rand(201, 1000)).T.astype(np.float32)).cuda()
        amp_dev = torch.empty_like(p_dev)
        wav_out = torch.empty(glay.sample_offsets[-1], dtype=torch.float32, device="cuda")
        out64 = torch.empty(glay.sample_offsets[-1], dtype=torch.float64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

        def gl_step():
            al._lib.check(gplan._lib.sc_power_to_amp_batch(gplan._h, p_dev.data_ptr(), glay.c_frame_offsets,
                                                           glay.c_frame_counts, GL_UTTS, 0.01, 1.0,
                                                           amp_dev.data_ptr(), st), "p2a")
            al.griffin_lim_device(gplan, amp_dev, ph_dev, glay, GL_ITERS, None, wav_out)
            al._lib.check(gplan._lib.sc_deemph_renorm_batch(gplan._h, wav_out.data_ptr(), glay.c_sample_offsets,
                                                            glay.c_sample_lengths, GL_UTTS, 0.97, 0.045,
                                                            out64.data_ptr(), st), "deemph")

        gl_step(); gl_step()
        barrier()
        gl_steps = max(2, min(args.steps, 5))
        tot = 0.0
        w0 = time.time()
        for _ in range(gl_steps):
            flush.zero_()                                   # L2 flush between timed steps (state < L2)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gl_step(); b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        windows.append((w0, time.time()))
        gl_ms = max_over_ranks(tot / gl_steps)
        gplan.profile(True)
        al.griffin_lim_device(gplan, amp_dev, ph_dev, glay, GL_ITERS, None, wav_out)
        pr = gplan.profile_read()
        gplan.profile(False)
        iter_ms = pr[1] / (GL_ITERS - 1)
        fr = GL_UTTS * 1000
        gl_audio = GL_UTTS * 80 * 999 / SR
        gl = {"workload": "configs[2]: 64 x (1000, 201) spectrograms, 200 iterations, fixed phase0, "
                          "mean_abs_amp_norm 0.045, realse 1.0 (test.py:148-156)",
              "value": world * gl_audio / (gl_ms * 1e-3), "unit": "audio-s/s", "ms_per_step": gl_ms,
              "ms_per_iteration": iter_ms, "l2": "flushed between timed steps; state stays L2-resident inside a step",
              "roofline": {"bound": "hbm", "kernel": "k_gl_iter_persist", "peak": peak, "unit": "GB/s",
                           "achieved": GL_BYTES_PRIMARY * fr / (iter_ms * 1e-3) / 1e9,
                           "frac": GL_BYTES_PRIMARY * fr / (iter_ms * 1e-3) / 1e9 / peak,
                           "definition": "4020 B/frame-iteration (complex64 spectrogram state, SURVEY.md §8(d))",
                           "strict_achieved": GL_BYTES_STRICT * fr / (iter_ms * 1e-3) / 1e9,
                           "strict_frac": GL_BYTES_STRICT * fr / (iter_ms * 1e-3) / 1e9 / peak,
                           "strict_definition": "1444 B/frame-iteration (waveform state: what the kernel moves)",
                           "traffic": traffic_tab.get("k_gl_iter_persist")}}

    # ---- long-form Griffin-Lim, configs[3]: one ~20 min spectrogram time-chunked over the ranks
    if gl is not None and args.gl_long:
        from speech_cloner_b200 import distributed as D
        T_long = 240001
        cg = D.ChunkedGriffinLim(T_long, 80, 400)
        f_lo, f_hi = cg.frame_range()
        reps = -(-(f_hi - f_lo) // 1000)
        amp_l = amp_dev[:1000].repeat(reps, 1)[: f_hi - f_lo].contiguous()      # tiled decoder-shaped magnitudes
        ph_l = ph_dev[:1000].repeat(reps, 1)[: f_hi - f_lo].contiguous()
        cg.run(amp_l, ph_l, 3)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); chunk = cg.run(amp_l, ph_l, GL_ITERS); b.record()
        barrier()
        long_ms = max_over_ranks(a.elapsed_time(b))
        gl["long_form"] = {"workload": "configs[3]: one (240001, 201) spectrogram (20 min), 200 iterations, time-chunked "
                                       f"over {world} rank(s), 480-sample halo exchange per iteration",
                           "ms_per_step": long_ms, "value": (80 * (T_long - 1) / SR) / (long_ms * 1e-3), "unit": "audio-s/s",
                           "scaling": "strong"}

    # ---- dataset-scale sweep, configs[4]: 10 h = 6000 x 3 s (TIMIT-shaped, gain 10) + 4500 x 4 s, sharded by frames
    sweep = None
    if args.sweep:
        from speech_cloner_b200 import distributed as D
        lens = [48000] * 6000 + [64000] * 4500
        mine = D.shard_by_frames(lens, world, 80)[rank]
        pool3 = synth.batch(5, 16, 3.0, ds_norm=(0.0, 10.0))
        pool4 = wavs[:16]
        slay = al.FrontendLayout([lens[i] for i in mine], 80)
        sdev = torch.empty(slay.total_samples, dtype=torch.float32, device="cuda")
        p3 = [torch.from_numpy(w).cuda() for w in pool3]
        p4 = [torch.from_numpy(w).cuda() for w in pool4]
        for k, (i, o) in enumerate(zip(mine, slay.sample_offsets)):
            src = p3[i % 16] if lens[i] == 48000 else p4[i % 16]
            sdev[o:o + lens[i]] = src                                     # 16 distinct utterances per shape, tiled
        sout = al.frontend_device(plan, sdev, slay)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            al.frontend_device(plan, sdev, slay, sout)
        b.record()
        barrier()
        sw_ms = max_over_ranks(a.elapsed_time(b) / 3)
        sweep = {"workload": "configs[4]: 10 h = 6000 x 3 s + 4500 x 4 s, utterances sharded by frame count (LPT) over "
                             f"{world} rank(s), device resident, 16 distinct synthetic utterances per shape tiled",
                 "ms_per_pass": sw_ms, "value": 36000.0 / (sw_ms * 1e-3), "unit": "audio-s/s", "scaling": "strong",
                 "frac_hbm": FE_BYTES_PER_FRAME * (6000 * 601 + 4500 * 801) / world / (sw_ms * 1e-3) / 1e9 / peak}
        del sdev, sout

    clocks = None
    if sampler:
        time.sleep(0.1)
        sampler.stop()
        clocks = sampler.summary(windows)
        clocks["sampled_over"] = "timed loops + a 1.2 s probe loop of the same front-end step + the Griffin-Lim steps"

    # ---- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v = cpu_frontend_single(wavs)
        cpu = {"value": v, "unit": "audio-s/s", "cores": 1, "kind": "port", "host_cores_available": os.cpu_count(),
               "sample": "all 256 utterances x 4 s once, sequential single-process loop like TIMIT_reader.py:169 "
                         "(oracle/audio_lib_oracle.py; librosa itself is not installable here)"}
        if gl is not None:
            P = Ps[0].cpu().numpy()
            np.random.seed(3000)
            ph = np.pi * np.random.rand(201, 1000)
            it = 10
            dt = cpu_gl_single(P, ph, it)
            gl["cpu_baseline"] = {"value": (80 * 999 / SR) / (dt * GL_ITERS / it), "unit": "audio-s/s", "cores": 1,
                                  "kind": "port", "sample": f"1 spectrogram x 5 s, {it} iterations timed, scaled linearly to 200"}

    if rank == 0:
        line = {
            "metric": "audio-seconds/second, front-end (calc_MFCC_input: STFT -> power dB, mel dB, MFCC+delta)",
            "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 FFT butterflies, f32 elsewhere" if args.precision == "fp64" else "f32",
            "data": "synthetic",
            "config": {"workload": "frontend configs[1]: 256 x 4 s ARCTIC-shaped utterances per GPU, hp/ds_dec_cfg_d.json "
                                   "(sr 16000, n_fft 400, hop 80, 80 mels, 40 MFCC + delta)",
                       "frames_per_step_per_gpu": frames, "fft_precision": args.precision,
                       "l2": "inputs+outputs (361.6 MB) larger than L2, no flush",
                       "parallelism": f"utterance shards, {world} rank(s), no data-path collective",
                       "cpu_binding_rank0": (f"{len(cpu_bind)} cores local to the GPU (NVML)" if cpu_bind else "none")},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "frontend_fp32_mode": alt, "griffin_lim": gl, "sweep_10h": sweep,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
