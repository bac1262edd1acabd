#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics, opcode histogram and the hottest SASS lines.
usage: python tools/ncu_summary.py report.ncu-rep [tiles_per_launch]"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
tiles = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w} [{units[i]}] = {[r[i] for r in data]}")
for i, h in enumerate(hdr):
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        v = [float(r[i]) for r in data]
        if max(v) > 0.15:
            print(f"{h.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', '')} = {v}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
        continue
    if cur is not None:
        cur["rows"].append(r)
b = blocks[0]
hdr, data = b["rows"][0], b["rows"][1:]
isrc, iex, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iex]) for r in data if r[iex].isdigit())
tots = sum(int(r[isamp]) for r in data if r[isamp].isdigit())
print("kernel", b["name"], "warp-instr", tot, "samples", tots, "per tile", tot / tiles if tiles else "")
c, s = Counter(), Counter()
for r in data:
    ex = int(r[iex]) if r[iex].isdigit() else 0
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
    op = m.group(2).split(".")[0] if m else "?"
    c[op] += ex
    s[op] += int(r[isamp]) if r[isamp].isdigit() else 0
for op, v in c.most_common(22):
    per = f"{v / tiles:8.1f}/tile" if tiles else ""
    print(f"  {op:10s} {v / tot * 100:6.2f}% {per}  samples {s[op] / max(tots, 1) * 100:5.1f}%")
print("hottest SASS by samples:")
top = sorted(data, key=lambda r: -(int(r[isamp]) if r[isamp].isdigit() else 0))[:25]
for r in top:
    print(f"  {int(r[isamp]) / max(tots, 1) * 100:5.2f}%  ex {r[iex]:>10}  {r[isrc][:90]}")
