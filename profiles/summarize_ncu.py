#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep brought back in gpurun_out/) into a small committed text file.

usage: python profiles/summarize_ncu.py gpurun_out/prof_X.ncu-rep profiles/r01_X.txt [kernel-note]
Reads the raw page (selected metrics), the details page (occupancy / issue / stall lines) and the
source page (opcode mix and the top source lines by executed instructions)."""
import collections
import csv
import io
import re
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
       "sass__inst_executed_global_loads", "sass__inst_executed_global_stores", "sass__inst_executed_local_loads",
       "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
       "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max"]
DETAIL = re.compile(r"Duration|Elapsed Cycles|Executed Ipc|Issue Slots Busy|No Eligible|Eligible Warps|Active Warps Per|"
                    r"Warp Cycles Per Issued|Avg. Active Threads|Registers Per|Theoretical Occ|Achieved Occ|L1/TEX Hit|"
                    r"L2 Hit|DRAM Throughput|Mem Busy|Max Bandwidth|stalled")


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    lines = [f"# ncu summary of {rep}  {note}", "# (ncu --set full --clock-control none; one launch; times are cold-cache, serialised)", ""]
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    if len(rows) >= 3:
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, zip(units, vals)))
            lines.append(f"## kernel: {d.get('Kernel Name', ('', '?'))[1]}")
            for k in RAW:
                if k in d:
                    lines.append(f"{k:70s} {d[k][1]:>18s} {d[k][0]}")
            lines.append("")
    lines.append("## details page (selected)")
    for ln in run(["-i", rep, "--page", "details"]).splitlines():
        if DETAIL.search(ln) and "OPT" not in ln:
            lines.append(ln.rstrip())
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    ops, tot, top = collections.Counter(), 0, []
    for r in src:
        if len(r) > 7 and r[0].strip().isdigit():
            try:
                top.append((int(r[7]), int(r[6]), int(r[0]), r[1].strip()[:100]))
            except ValueError:
                pass
        elif len(r) > 7 and r[2].startswith("0x"):
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3])
            try:
                n = int(r[7])
            except ValueError:
                continue
            if m:
                ops[m.group(2)] += n
                tot += n
    if tot:
        lines += ["", "## opcode mix (warp-level instructions executed)"]
        for o, n in ops.most_common(16):
            lines.append(f"{o:10s} {100.0 * n / tot:6.2f} %")
    if top:
        top.sort(reverse=True)
        t = sum(x[0] for x in top)
        lines += ["", "## top source lines by executed instructions (share, stall samples, line, text)"]
        for n, s, ln, txt in top[:25]:
            lines.append(f"{100.0 * n / t:5.1f} %  {s:8d}  L{ln:<4d} {txt}")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
