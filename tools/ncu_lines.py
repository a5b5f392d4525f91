#!/usr/bin/env python3
"""Per source line of one kernel in an .ncu-rep: warp instructions executed, average active threads, shared-memory
wavefronts and stall samples (ncu --page source, cuda,sass view aggregated by line).
usage: tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top N]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
agg = {}
cur_line = None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    ix = {h: i for i, h in enumerate(hdr)}
    if r[0].strip():
        cur_line = (int(r[0]), r[1].strip())
    if cur_line is None:
        continue
    def f(name):
        try:
            return float(r[ix[name]])
        except (ValueError, KeyError, IndexError):
            return 0.0
    a = agg.setdefault(cur_line, [0, 0, 0, 0, 0])
    a[0] += f("Instructions Executed")
    a[1] += f("Thread Instructions Executed")
    a[2] += f("L1 Wavefronts Shared")
    a[3] += f("# Samples")
    a[4] += f("L1 Wavefronts Shared Ideal")
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[3] for a in agg.values()) or 1
print(f"total warp instructions {tot / 1e9:.3f} G, samples {tots:.0f}")
print(f"{'line':>5} {'inst %':>7} {'samp %':>7} {'thr/inst':>8} {'smem wf (M)':>11} {'ideal':>8}  source")
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln:5d} {100 * a[0] / tot:7.2f} {100 * a[3] / tots:7.2f} {a[1] / a[0] if a[0] else 0:8.1f} {a[2] / 1e6:11.1f} {a[4] / 1e6:8.1f}  {src[:110]}")
