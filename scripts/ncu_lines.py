#!/usr/bin/env python
"""Per-source-line summary of an `ncu --page source --csv` export: instructions, shared-memory wavefronts, stall samples.

    python scripts/ncu_lines.py <src.csv> [top_n] [sort: inst|wf|stall]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
key = sys.argv[3] if len(sys.argv) > 3 else "inst"
cur, hdr, agg = None, None, {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < 20 or r[0] == '':
        continue
    def g(name):
        try: return int(r[hdr.index(name)])
        except Exception: return 0
    agg[(cur, int(r[0]))] = dict(inst=g('Instructions Executed'), wf=g('L1 Wavefronts Shared'), exc=g('L1 Wavefronts Shared Excessive'),
                                 stall=g('Warp Stall Sampling (All Samples)'), src=r[1].strip()[:100])
ti = sum(v['inst'] for v in agg.values()); tw = sum(v['wf'] for v in agg.values()); ts = sum(v['stall'] for v in agg.values())
print(f"total inst {ti}  smem wavefronts {tw}  stall samples {ts}")
for k, v in sorted(agg.items(), key=lambda x: -x[1][key])[:top]:
    print(f"{k[0]:16s} {k[1]:4d} inst {100*v['inst']/ti:5.1f}% wf {100*v['wf']/max(tw,1):5.1f}% (exc {v['exc']:9d}) stall {100*v['stall']/max(ts,1):5.1f}%  {v['src']}")
