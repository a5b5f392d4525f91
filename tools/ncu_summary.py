#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full ... --import-source on) into markdown for profiles/.
usage: tools/ncu_summary.py <report.ncu-rep> <out.md> [title]"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, out_md = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = ncu(rep, "raw")
    hdr, units = raw[0], raw[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    lines = [f"# {title}", "", f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`); one section per captured launch.", ""]
    seen = {}
    for r in raw[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        seen[name] = seen.get(name, 0) + 1
        if seen[name] > 1:
            continue
        lines += [f"## `{name}`", "", "| metric | value | unit |", "|---|---|---|"]
        for m in METRICS:
            if m in idx and r[idx[m]] != "":
                lines.append(f"| {m} | {r[idx[m]]} | {units[idx[m]]} |")
        st = sorted(((float(r[idx[h]].replace(",", "")), h) for h in stall_cols if r[idx[h]]), reverse=True)[:8]
        lines += ["", "Warp stall reasons (warps stalled per issue-active cycle):", ""]
        for v, h in st:
            lines.append(f"- {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}")
        lines.append("")
    # per-instruction hot spots of each kernel
    src = ncu(rep, "source")
    starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
    done = set()
    for k, s0 in enumerate(starts):
        name = src[s0][1].split("(")[0]
        if name in done:
            continue
        done.add(name)
        e0 = starts[k + 1] if k + 1 < len(starts) else len(src)
        h = src[s0 + 1]
        try:
            ai, si, ei = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
        except ValueError:
            continue
        data = [(int(r[si] or 0), int(r[ei] or 0), r[ai].strip()) for r in src[s0 + 2:e0] if len(r) > ei]
        tot = sum(d[0] for d in data) or 1
        lines += [f"### `{name}`: top SASS instructions by stall samples ({tot} samples, {sum(d[1] for d in data) / 1e9:.3f} G warp instructions)", "",
                  "| samples | share | executed | instruction |", "|---|---|---|---|"]
        for s, e, a in sorted(data, reverse=True)[:14]:
            lines.append(f"| {s} | {100 * s / tot:.1f}% | {e} | `{a}` |")
        lines.append("")
    open(out_md, "w").write("\n".join(lines) + "\n")
    print("wrote", out_md)


if __name__ == "__main__":
    main()
