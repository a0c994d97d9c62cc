#!/usr/bin/env python
"""bench.py — audio-seconds/second of the speech-cloner DSP hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--precision fp64|fp32] [--no-gl] [--no-gl-long] [--no-sweep] [--no-cpu] [--no-e2e]

A "step" is one pass of the front-end (calc_MFCC_input, /root/reference/audio_lib.py:89-244) over
BASELINE.json configs[1]: 256 synthetic ARCTIC-shaped utterances x 4 s at the hp/*.json settings
(205 056 frames, 65.5 MB in, 296.1 MB out).  One JSON line is printed by rank 0:

  value        whole-job audio-s/s with the waveforms already resident in HBM (CUDA events, max over ranks)
  e2e          the same through the host-buffer pipeline: pinned H2D of every waveform and pinned D2H of the
               three feature arrays inside the timed region; pcie_* = the same bytes moved by bare pinned copies
               (both directions at once, all ranks at once) in this run, i.e. the ceiling e2e can reach
  roofline     achieved algorithmic GB/s (1 764 B/frame) over the summed kernel time of a
               step, against MEASURED_PEAKS.json hbm_gbs; per-kernel shares from CUDA events in the library;
               traffic = DRAM bytes from the committed ncu capture, reported only while the kernel sources still
               hash to what was captured
  cpu_baseline the CPU oracle (restated reference path) on this box's host cores, bounded sample
  griffin_lim  configs[2]: 64 spectrograms x 5 s, 200 iterations (test.py:87), device-resident and `e2e` through
               from_power_to_wav_batch with host arrays; long_form = configs[3]: one 20 min spectrogram time-chunked
               over the ranks, prologue + 200 iterations + distributed epilogue + final gather inside the timed region
  sweep_10h    configs[4]: 10 h of audio, utterance-sharded, with the final gather of the feature buffers timed
  parity_selfcheck   sharded == single and chunked == unchunked (bit-identical) on this run's ranks

Under torchrun every rank featurises its own copy of the batch (weak scaling, no data-path collective: utterances
are independent); configs[3] and configs[4] are strong-scaling legs.  `--impl reference` times the
CPU oracle with all host cores on the same config (rank 0 only) and prints the same metric string.
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

# ONE metric / workload string for both arms: the driver divides the arms only when they match
METRIC = "audio-seconds/second, front-end (calc_MFCC_input)"
UNIT = "audio-s/s"
WORKLOAD = ("frontend configs[1]: 256 x 4 s ARCTIC-shaped utterances per GPU, hp/ds_dec_cfg_d.json "
            "(sr 16000, n_fft 400, hop 80, 80 mels, 40 MFCC + delta)")

N_UTTS, SECONDS, SR = 256, 4.0, 16000
GL_UTTS, GL_SECONDS, GL_ITERS = 64, 5.0, 200
T_LONG = 240001                                            # configs[3]: 20 min at hop 80
FE_BYTES_PER_FRAME = 4 * (80 + 201 + 80 + 80)            # wav in + three outputs
GL_BYTES_PRIMARY = 4 * 201 + 8 * 201 + 8 * 201           # complex64 spectrogram state
GL_BYTES_STRICT = 4 * 201 + 2 * 4 * 80                   # waveform state (what the kernel moves)
GL_KW = dict(P_dB_norm_factor=0.01, pre_emphasis=0.97, hop_length=80, win_length=400, mean_abs_amp_norm=0.045,
             realse=1.0)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_traffic.json")


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_source_hash() -> str:
    """sha256 over the kernel sources and the header: pins profiles/r02_traffic.json to the code it was captured from."""
    from speech_cloner_b200 import build
    return build.source_hash()


def load_traffic():
    """(table, note): the committed ncu DRAM-byte table if it still belongs to the current kernel sources, else {}."""
    try:
        with open(TRAFFIC_FILE) as f:
            tab = json.load(f)
    except Exception:
        return {}, "no committed ncu traffic capture for this round (profiles/r02_traffic.json)"
    want, have = tab.get("_src_sha256"), kernel_source_hash()
    if want != have:
        return {}, (f"profiles/r02_traffic.json was captured from other kernel sources (sha256 {str(want)[:12]} != "
                    f"{have[:12]}): stale, not reported")
    return tab, f"DRAM read+write bytes per launch, ncu --set full, profiles/r02_traffic.json (source sha256 {have[:12]})"


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
    oracle.from_power_to_wav(P, n_iter=iters, verbose=False, phase0=phase0, **GL_KW)
    return time.perf_counter() - t


def run_reference(args):
    """`--impl reference`: the reference's CPU path (oracle restatement; librosa is not installable) with
    every host core, one utterance per task, BLAS threads pinned to 1, on ALL 256 utterances of the same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from speech_cloner_b200 import synth
    cores = os.cpu_count() or 1
    wavs = synth.batch(2, N_UTTS, SECONDS)
    audio_s = N_UTTS * SECONDS
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_oracle_utt, wavs, chunksize=1)
        t = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_oracle_utt, wavs, chunksize=1)
        dt = (time.perf_counter() - t) / args.steps
    val = audio_s / dt
    sample = f"all {N_UTTS} utterances x {SECONDS:g} s per step (the full config), multiprocessing.Pool({cores})"
    emit({
        "impl": "reference", "metric": METRIC, "value": val,
        "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "same_config": True, "sample": sample,
                   "note": "the reference is one single-process CPU job whatever --gpus says; value is its whole-job rate"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------ GPU legs
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--no-gl", action="store_true", help="skip every Griffin-Lim leg")
    ap.add_argument("--no-gl-long", action="store_true", help="skip the chapter-length chunked Griffin-Lim (configs[3])")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 10 h featurization sweep (configs[4])")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs (profiling runs)")
    ap.add_argument("--no-probe", action="store_true", help="skip the 1.2 s clock probe loop (profiling runs)")
    ap.add_argument("--no-selfcheck", action="store_true", help="skip the distributed parity self-check")
    ap.add_argument("--gl-long", action="store_true", help=argparse.SUPPRESS)    # round-1 spelling: now the default
    ap.add_argument("--sweep", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from speech_cloner_b200 import audio_lib as al
    from speech_cloner_b200 import distributed as D
    from speech_cloner_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    bind = {"action": "not attempted", "why": "single rank"}
    if world > 1:
        # one process per GPU: keep the rank (and the pinned buffers it is about to allocate) on the GPU's NUMA node
        bind = D.bind_report(local)
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

    def timed(fn, reps=1):
        """max over ranks of the device time of fn() (CUDA events on the current stream, barrier both sides)."""
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        res = None
        for _ in range(reps):
            res = fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b) / reps), res

    peak, peak_src = hbm_peak()
    traffic_tab, traffic_note = load_traffic()
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
    names = ["gain: k_abs_pairwise4 + k_gain_finalize (k_fe_setup only when the layout changes)", "pass A: k_fe_pass_a_ws",
             "pass B: k_fe_pass_b3 (MFCC[0,0] inside)"]
    tkeys = ["gain", "pass_a", "pass_b"]
    top = int(np.argmax(prof))
    fe_bytes = FE_BYTES_PER_FRAME * frames
    achieved = fe_bytes / (kern_ms * 1e-3) / 1e9
    kbytes = [320, 1444, 2568]
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": (sum(traffic_tab.get(k, float("nan")) for k in tkeys) if traffic_tab else None),
        "traffic_note": traffic_note + "; algorithmic bytes of the step: %d" % fe_bytes,
        "kernel_traffic": traffic_tab.get(tkeys[top]),
        "peak_source": peak_src,
        "definition": "1764 B/frame x frames of one step / summed device time of the step's kernels",
        "kernel": names[top],
        # the dominant kernel against its own algorithmic bytes: gain reads the audio once
        # (320 B/frame), pass A reads it again and writes the raw dB tiles (320 + 1124), pass B re-reads and rewrites
        # (1124 + 1444)
        "kernel_algorithmic_bytes_per_frame": kbytes[top],
        "kernel_achieved": kbytes[top] * frames / (float(prof[top]) * 1e-3) / 1e9,
        "kernel_frac": kbytes[top] * frames / (float(prof[top]) * 1e-3) / 1e9 / peak,
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
        ms32, _ = timed(lambda: al.frontend_device(plan32, wav_dev, lay, out), args.steps)
        alt = {"fft_precision": "fp32", "ms_per_step": ms32, "value": world * audio_s / (ms32 * 1e-3), "unit": UNIT,
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
        # PCIe ceiling measured in this run: the step's H2D and D2H bytes as bare pinned copies on two streams, both
        # directions at once, every rank at once (ranks of one node share the host's PCIe-to-memory bandwidth)
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

        def bare_copies():
            with torch.cuda.stream(s_up):
                pipe.wav_dev.copy_(pipe.wav_host, non_blocking=True)
            with torch.cuda.stream(s_dn):
                for h, d in zip(pipe.out_host, pipe.out_dev):
                    h.copy_(d, non_blocking=True)
        for _ in range(2):
            bare_copies()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            bare_copies()
            s_up.synchronize(); s_dn.synchronize()
        pcie_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        e2e = {"value": world * audio_s / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": int(pipe.h2d_bytes), "d2h_bytes_per_step": int(pipe.d2h_bytes),
               "path": "FrontendPipeline.run(): 8 utterance chunks over 3 streams, pinned host buffers",
               "pcie_ms_per_step": pcie_s * 1e3,
               "pcie_ceiling_gbs": world * (pipe.h2d_bytes + pipe.d2h_bytes) / pcie_s / 1e9,
               "pcie_ceiling_value": world * audio_s / pcie_s,
               "frac_of_pcie": pcie_s / e2e_s,
               "pcie_note": f"the same {pipe.h2d_bytes + pipe.d2h_bytes} bytes per rank as bare pinned copies, both directions and "
                            f"all {world} rank(s) concurrently, measured in this run: e2e cannot exceed pcie_ceiling_value"}

    # ---- Griffin-Lim, configs[2]
    gl = None
    Ps_host = phs_host = None
    if not args.no_gl:
        gl_wavs = synth.batch(3 + 10 * rank, GL_UTTS, GL_SECONDS)
        feats = al.calc_MFCC_input_batch(gl_wavs, return_device=True, **hp)
        Ps = [f[2][:1000].contiguous() for f in feats]                     # (1000, 201) decoder-shaped
        glay = al._GlLayout([1000] * GL_UTTS, 80)
        gplan = al.DspPlan.get(n_fft=400, win_length=400, hop_length=80)
        p_dev = torch.zeros((glay.frame_offsets[-1], 201), dtype=torch.float32, device="cuda")
        ph_dev = torch.zeros_like(p_dev)
        phs_host = []
        for i, (P, o) in enumerate(zip(Ps, glay.frame_offsets)):
            p_dev[o:o + 1000] = P
            np.random.seed(3000 + i)
            phs_host.append(np.pi * np.random.rand(201, 1000))                # float64 (bins, T), like audio_lib.py:255
            ph_dev[o:o + 1000] = torch.from_numpy(phs_host[-1].T.astype(np.float32)).cuda()
        Ps_host = [P.cpu().numpy() for P in Ps]
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
              "value": world * gl_audio / (gl_ms * 1e-3), "unit": UNIT, "ms_per_step": gl_ms,
              "ms_per_iteration": iter_ms, "l2": "flushed between timed steps; state stays L2-resident inside a step",
              "roofline": {"bound": "hbm", "kernel": "k_gl_iter_persist", "peak": peak, "unit": "GB/s",
                           "achieved": GL_BYTES_PRIMARY * fr / (iter_ms * 1e-3) / 1e9,
                           "frac": GL_BYTES_PRIMARY * fr / (iter_ms * 1e-3) / 1e9 / peak,
                           "definition": "4020 B/frame-iteration (complex64 spectrogram state, SURVEY.md §8(d)); bookkeeping "
                                         "figure: the kernel keeps the waveform state, see strict_*",
                           "strict_achieved": GL_BYTES_STRICT * fr / (iter_ms * 1e-3) / 1e9,
                           "strict_frac": GL_BYTES_STRICT * fr / (iter_ms * 1e-3) / 1e9 / peak,
                           "strict_definition": "1444 B/frame-iteration (waveform state: what the kernel moves)",
                           "traffic": traffic_tab.get("gl_iter")}}
        # ---- the same through the public API with HOST arrays (from_power_to_wav_batch: pinned staging, one H2D per
        # array kind, device transpose of the float64 phases, pinned D2H of the float64 waveforms)
        if not args.no_e2e:
            def gl_public():
                return al.from_power_to_wav_batch(Ps_host, n_iter=GL_ITERS, verbose=False, phase0s=phs_host, n_fft=None, **GL_KW)
            res = gl_public()
            res = gl_public()          # second warm call while the first result is alive: the pinned pool reaches its steady two blocks
            barrier()
            reps = 3
            t0 = time.perf_counter()
            for _ in range(reps):
                res = gl_public()
            gl_e2e_s = max_over_ranks((time.perf_counter() - t0) / reps)
            gl["e2e"] = {"value": world * gl_audio / gl_e2e_s, "unit": UNIT, "ms_per_step": gl_e2e_s * 1e3,
                         "h2d_bytes_per_step": int(sum(P.nbytes for P in Ps_host) + sum(p.nbytes for p in phs_host)),
                         "d2h_bytes_per_step": int(sum(r.nbytes for r in res)),
                         "path": "audio_lib.from_power_to_wav_batch(list of host (1000, 201) float32 maps, host float64 phases) -> "
                                 "host float64 waveforms"}

    # ---- long-form Griffin-Lim, configs[3]: one ~20 min spectrogram time-chunked over the ranks.  Timed region =
    # prologue (power -> amplitude) + 200 iterations with a halo exchange every k + distributed epilogue + final gather
    long_full = None
    if gl is not None and not args.no_gl_long:
        # ONE signal for all ranks: rank 0's first decoder-shaped map and phase, tiled to 20 min (the per-rank batches
        # above use different seeds per rank, so they cannot be mixed into one spectrogram)
        p1000, ph1000 = p_dev[:1000].clone(), ph_dev[:1000].clone()
        if world > 1:
            dist.broadcast(p1000, src=0)
            dist.broadcast(ph1000, src=0)
        cg = D.ChunkedGriffinLim(T_LONG, 80, 400, steps_per_exchange=40)
        f_lo, f_hi = cg.frame_range(n_iters=GL_ITERS)
        idx = torch.arange(f_lo, f_hi, device="cuda") % 1000
        p_l = p1000[idx].contiguous()                                          # tiled decoder-shaped power-dB rows
        ph_l = ph1000[idx].contiguous()

        def long_step():
            y = cg.from_power_to_wav(p_l, ph_l, n_iter=GL_ITERS, **{k: v for k, v in GL_KW.items()
                                                                    if k not in ("hop_length", "win_length")})
            return cg.gather(y, dst=None)
        long_step()
        long_ms, long_full = timed(long_step)
        n_exch = -(-GL_ITERS // cg._k_for(GL_ITERS)) - 1 if world > 1 else 0
        gl["long_form"] = {"workload": f"configs[3]: one ({T_LONG}, 201) spectrogram (20 min), 200 iterations, time-chunked over "
                                       f"{world} rank(s); timed: sc_p2a_chunk_* + {GL_ITERS} iterations ({n_exch} halo exchanges of "
                                       f"{cg._k_for(GL_ITERS) * cg.halo} samples) + distributed de-emphasis / renorm + all_gather of the "
                                       "float64 waveform",
                           "ms_per_step": long_ms, "value": (80 * (T_LONG - 1) / SR) / (long_ms * 1e-3), "unit": UNIT,
                           "scaling": "strong", "steps_per_exchange": cg._k_for(GL_ITERS), "halo_samples_per_iteration": cg.halo}

    # ---- dataset-scale sweep, configs[4]: 10 h = 6000 x 3 s (TIMIT-shaped, gain 10) + 4500 x 4 s, sharded by frames,
    # compute + the one final gather of the three feature buffers
    sweep = None
    if not args.no_sweep:
        lens = [48000] * 6000 + [64000] * 4500
        shards = D.shard_by_frames(lens, world, 80)
        mine = shards[rank]
        pool3 = synth.batch(5, 16, 3.0, ds_norm=(0.0, 10.0))
        pool4 = wavs[:16]
        slay = al.FrontendLayout([lens[i] for i in mine], 80)
        rows = [al.FrontendLayout([lens[i] for i in s], 80).total_frames for s in shards]
        sdev = torch.empty(slay.total_samples, dtype=torch.float32, device="cuda")
        p3 = [torch.from_numpy(w).cuda() for w in pool3]
        p4 = [torch.from_numpy(w).cuda() for w in pool4]
        for k, (i, o) in enumerate(zip(mine, slay.sample_offsets)):
            src = p3[i % 16] if lens[i] == 48000 else p4[i % 16]
            sdev[o:o + lens[i]] = src                                     # 16 distinct utterances per shape, tiled
        # outputs padded to the largest shard so the gather needs no staging copy
        widths = (plan.mfcc_width, plan.n_mels, plan.n_bins)
        sout = tuple(torch.empty((max(rows), w), dtype=torch.float32, device="cuda") for w in widths)
        gathered = tuple(torch.empty((world * max(rows), w), dtype=torch.float32, device="cuda") for w in widths) if world > 1 else None

        def sweep_compute():
            al.frontend_device(plan, sdev, slay, sout)

        def sweep_gather():
            if world > 1:
                for g, s in zip(gathered, sout):
                    dist.all_gather_into_tensor(g, s)

        def sweep_all():
            sweep_compute(); sweep_gather()
        sweep_all()
        sw_c, _ = timed(sweep_compute, 3)
        sw_all, _ = timed(sweep_all, 3)
        gbytes = 4 * sum(widths) * max(rows) * (world - 1)                # bytes each rank receives
        sweep = {"workload": "configs[4]: 10 h = 6000 x 3 s + 4500 x 4 s, utterances sharded by frame count (LPT) over "
                             f"{world} rank(s), device resident, 16 distinct synthetic utterances per shape tiled; timed: front-end "
                             "of the shard + all_gather of the three packed feature buffers (the one collective of the path)",
                 "ms_per_pass": sw_all, "ms_compute": sw_c, "ms_gather": sw_all - sw_c,
                 "value": 36000.0 / (sw_all * 1e-3), "value_compute_only": 36000.0 / (sw_c * 1e-3), "unit": UNIT,
                 "scaling": "strong", "gather_bytes_received_per_rank": int(gbytes),
                 "gather_gbs_per_rank": (gbytes / ((sw_all - sw_c) * 1e-3) / 1e9) if world > 1 and sw_all > sw_c else None,
                 "frac_hbm_compute": FE_BYTES_PER_FRAME * (6000 * 601 + 4500 * 801) / world / (sw_c * 1e-3) / 1e9 / peak}
        del sdev, sout, gathered

    # ---- window sampler on the device-resident features of the step (every rank its own cache)
    wsamp = None
    try:
        from speech_cloner_b200 import dataset_cache as dc
        cache = dc.DeviceSpecCache.from_device(out[0], out[1], out[2], lay)
        n_t, n_w = 400, 2048                                          # 2 048 windows of 400 frames = 1.18 GB per launch
        rs = np.random.RandomState(5)
        u = rs.randint(0, N_UTTS, size=n_w)
        first = cache.frame_offsets[u] + rs.randint(0, np.asarray(cache.spec_len)[u] - n_t)
        valid = np.full(n_w, n_t, np.int32)
        ws_out = cache.gather(dc.DeviceSpecCache.FEATURES, first, valid, n_t)
        ws_ms, _ = timed(lambda: cache.gather(dc.DeviceSpecCache.FEATURES, first, valid, n_t, out=ws_out) and None, 5)
        ws_bytes = 2 * 4 * n_w * n_t * sum(out_i.shape[1] for out_i in out)
        np.random.seed(3)
        kw_s = dict(batch_size=32, prop_val=0.0, verbose=False)
        for _ in dc.spec_window_sampler(cache, range(N_UTTS), n_t, n_epochs=2, **kw_s):      # warm the allocators
            pass
        barrier()
        t0 = time.perf_counter()
        nb = sum(1 for _ in dc.spec_window_sampler(cache, range(N_UTTS), n_t, n_epochs=16, **kw_s))
        torch.cuda.synchronize()
        ws_batch_us = 1e6 * (time.perf_counter() - t0) / max(nb, 1)
        wsamp = {"workload": f"{n_w} random windows x {n_t} frames x (80 + 80 + 201) float32 out of the step's packed outputs, "
                             "one sc_window_gather launch into preallocated tensors (incl. the pinned index upload of DeviceSpecCache.gather)",
                 "ms": ws_ms, "achieved": ws_bytes / (ws_ms * 1e-3) / 1e9, "unit": "GB/s", "peak": peak,
                 "frac": ws_bytes / (ws_ms * 1e-3) / 1e9 / peak,
                 "bytes": "windows read + written; the 296 MB source is read ~4 times over, so most reads hit L2 and the DRAM side is "
                          "the 1.18 GB written (a 1.18 GB cache read once per window runs at 4.8 TB/s, profiles/r02_sampler_bench.json)",
                 "us_per_32_window_batch_through_spec_window_sampler": ws_batch_us, "ranks": world}
        del cache, ws_out
    except Exception as e:                                            # never lose the bench line to a side leg
        wsamp = {"error": f"{type(e).__name__}: {e}"}

    # ---- distributed parity self-check on this run's ranks (the 1-GPU test run cannot see these)
    selfcheck = None
    if not args.no_selfcheck and gl is not None:
        selfcheck = {}
        try:
            if long_full is None:
                p1000, ph1000 = p_dev[:1000].clone(), ph_dev[:1000].clone()
                if world > 1:
                    dist.broadcast(p1000, src=0)
                    dist.broadcast(ph1000, src=0)
            # (1) sharded == single, bit for bit
            sw = [synth.utterance(900 + i, s) for i, s in enumerate([1.0, 0.4, 2.0, 0.7, 1.3, 0.2, 0.9, 3.0])]
            got = D.featurize_sharded(sw, gather=True, return_device=True, **hp)
            want = al.calc_MFCC_input_batch(sw, return_device=True, **hp)
            ok1 = all(torch.equal(a, b) for g, w in zip(got, want) for a, b in zip(g, w))
            selfcheck["sharded_equals_single"] = bool(ok1)
            # (2) chunked == unchunked at T = 2001 (realse 1.2 exercises the distributed prologue) ...
            def unchunked(T, n_iter, realse):
                ii = torch.arange(T, device="cuda") % 1000
                kw = dict(GL_KW); kw["realse"] = realse
                return al.from_power_to_wav_batch([p1000[ii]], n_iter=n_iter, verbose=False, n_fft=None,
                                                  phase0s=[ph1000[ii].t()], return_device=True, **kw)[0]

            def chunked(T, n_iter, realse, k):
                c = D.ChunkedGriffinLim(T, 80, 400, steps_per_exchange=k)
                lo_f, hi_f = c.frame_range(n_iters=n_iter)
                ii = torch.arange(lo_f, hi_f, device="cuda") % 1000
                y = c.from_power_to_wav(p1000[ii].contiguous(), ph1000[ii].contiguous(), P_dB_norm_factor=0.01,
                                        pre_emphasis=0.97, mean_abs_amp_norm=0.045, n_iter=n_iter, realse=realse)
                return c.gather(y, dst=None)
            ok2 = bool(torch.equal(chunked(2001, 25, 1.2, 4), unchunked(2001, 25, 1.2)))
            selfcheck["chunked_equals_unchunked_T2001"] = ok2
            # ... and at the full T = 240 001 on the waveform the timed long-form step produced
            ok3 = None
            if long_full is not None:
                ok3 = bool(torch.equal(long_full, unchunked(T_LONG, GL_ITERS, 1.0)))
                selfcheck["chunked_equals_unchunked_T240001"] = ok3
            flags = torch.tensor([int(ok1), int(ok2), int(ok3 is not False)], device="cuda")
            if world > 1:
                dist.all_reduce(flags, op=dist.ReduceOp.MIN)
            selfcheck["all_ranks"] = bool(flags.min().item() == 1)
            selfcheck["result"] = "ok" if selfcheck["all_ranks"] else "MISMATCH"
        except Exception as e:                                   # never lose the bench line to a check
            selfcheck["result"] = f"error: {type(e).__name__}: {e}"
        selfcheck["ranks"] = world
        selfcheck["what"] = ("bit-identity (torch.equal) of featurize_sharded vs one batch call, and of the time-chunked "
                             "from_power_to_wav (NCCL halo exchanges, distributed prologue / epilogue, gather) vs the single-GPU call")

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
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "host_cores_available": os.cpu_count(),
               "sample": "all 256 utterances x 4 s once, sequential single-process loop like TIMIT_reader.py:169 "
                         "(oracle/audio_lib_oracle.py; librosa itself is not installable here)"}
        if gl is not None:
            it = 10
            dt = cpu_gl_single(Ps_host[0], phs_host[0], it)
            gl["cpu_baseline"] = {"value": (80 * 999 / SR) / (dt * GL_ITERS / it), "unit": UNIT, "cores": 1,
                                  "kind": "port", "sample": f"1 spectrogram x 5 s, {it} iterations timed, scaled linearly to 200"}

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 FFT butterflies, f32 elsewhere" if args.precision == "fp64" else "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "same_config": True,
                       "frames_per_step_per_gpu": frames, "fft_precision": args.precision,
                       "l2": "inputs+outputs (361.6 MB) larger than L2, no flush",
                       "parallelism": f"utterance shards, {world} rank(s), no data-path collective",
                       "cpu_binding_rank0": f"{bind['action']}: {bind['why']}"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "frontend_fp32_mode": alt, "griffin_lim": gl, "sweep_10h": sweep, "window_sampler": wsamp, "parity_selfcheck": selfcheck,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
