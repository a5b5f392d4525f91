#!/usr/bin/env python3
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X) into markdown for profiles/.
usage: tools/launch_summary.py <launches.csv> <out.md> "<title>" "<command>" """
import csv
import sys


def main():
    src, out, title, cmd = sys.argv[1:5]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        t = float(r[v].replace(",", ""))
        t = t / 1000.0 if r[u] in ("ns", "nsecond") else t * (1000.0 if r[u] in ("ms", "msecond") else 1.0)
        name = r[k].split("(")[0][:72]
        n, s = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, s + t)
    total = sum(s for _, s in agg.values())
    lines = [f"# {title}", "", f"Command: `{cmd}`",
             "(per-launch times are cold-cache and serialised under ncu: compare shares, not absolutes; torch's own "
             "kernels are the bench's verification compares)", "",
             "| kernel | launches | avg us | total us | share |", "|---|---|---|---|---|"]
    for name, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{name}` | {n} | {s / n:.1f} | {s:.1f} | {100 * s / total:.2f}% |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
