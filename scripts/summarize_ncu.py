#!/usr/bin/env python
"""Turn ncu captures into the tracked summaries under profiles/.

    python scripts/summarize_ncu.py launches <launches.csv> <out.md>      # --metrics gpu__time_duration.sum list
    python scripts/summarize_ncu.py full <prof.ncu-rep> <out.md>          # --set full capture (needs ncu here)
    python scripts/summarize_ncu.py traffic <prof.ncu-rep> <out.json>     # DRAM bytes per launch for bench.py's roofline.traffic
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("launch__occupancy_limit_registers", "CTAs/SM (reg limit)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: mio"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[rows.index(hdr) + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        agg.setdefault(r[ki].split("(")[0], []).append(v)
    tot = sum(sum(v) for k, v in agg.items() if "scdsp" in k)
    with open(dst, "w") as f:
        f.write("| kernel | launches | mean us | min us | max us | share of scdsp time |\n|---|---:|---:|---:|---:|---:|\n")
        for k, v in agg.items():
            share = f"{100 * sum(v) / tot:.1f} %" if "scdsp" in k else "-"
            f.write(f"| `{k[:70]}` | {len(v)} | {sum(v) / len(v):.1f} | {min(v):.1f} | {max(v):.1f} | {share} |\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    seen = collections.OrderedDict()
    for r in rows[2:]:
        seen[r[ki].split("(")[0]] = r          # keep the last (warm) launch of each kernel
    with open(dst, "w") as f:
        names = list(seen)
        f.write("| metric | " + " | ".join(f"`{n[-38:]}`" for n in names) + " |\n|---|" + "---:|" * len(names) + "\n")
        for key, label in KEYS:
            if key not in hdr:
                continue
            i = hdr.index(key)
            cells = []
            for n in names:
                v = seen[n][i]
                try:
                    x = float(v.replace(",", ""))
                    v = f"{x:,.2f}" if x < 1000 else f"{x:,.0f}"
                except ValueError:
                    pass
                cells.append(f"{v} {units[i]}".strip())
            f.write(f"| {label} | " + " | ".join(cells) + " |\n")


GROUPS = [   # bench.py roofline groups -> kernels of one front-end step
    ("gain", ("k_fe_setup", "k_abs_pairwise", "k_gain_finalize")),
    ("pass_a", ("k_fe_pass_a",)),
    ("pass_b", ("k_fe_c00", "k_fe_pass_b")),
    ("gl_iter", ("k_gl_iter_persist",)),
]


def traffic(src, dst):
    """DRAM read + write bytes per launch (last = warm launch of every kernel), summed per bench group, and pinned to
    the kernel sources they were captured from (bench.py reports them only while the hash still matches)."""
    import json
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    seen = collections.OrderedDict()
    for r in rows[2:]:
        seen[r[ki].split("(")[0].split("::")[-1]] = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
    from speech_cloner_b200 import build
    sha = build.source_hash()
    res = {"_source": f"{src}: dram__bytes_read.sum + dram__bytes_write.sum per launch (last launch of each kernel), "
                      "ncu --set full, config-2 front-end shape (256 x 4 s) and config-3 Griffin-Lim shape",
           "_src_sha256": sha,
           "_kernels": seen}
    for name, pats in GROUPS:
        res[name] = sum(v for k, v in seen.items() if any(p_ in k for p_ in pats))
    with open(dst, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
