"""Executed warp instructions and stall samples per CUDA source line, from the SASS rows of an
.ncu-rep (each SASS row is attributed to the last CUDA line marker above it).
usage: python tools/ncu_sass_lines.py report.ncu-rep units_per_launch [top]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep, units = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
iaddr = hdr.index("Address")
ex, samp, text = defaultdict(int), defaultdict(int), {}
fname, line = None, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) <= isamp:
        continue
    if r[0].strip().isdigit():  # a CUDA line row (its counters are the sum of its SASS rows: skip them)
        line = (fname, int(r[0]))
        text[line] = r[1]
        continue
    if line is None or not r[iaddr].strip():
        continue
    if r[isamp].strip().isdigit():
        samp[line] += int(r[isamp])
    if r[iex].strip().isdigit():
        ex[line] += int(r[iex])
tot, ts = sum(ex.values()), sum(samp.values()) or 1
print(f"warp instructions per unit: {tot / units:.1f}; samples {ts}")
for k in sorted(ex, key=lambda k: -ex[k])[:top]:
    print(f"{ex[k] / units:8.1f}/unit {100 * ex[k] / tot:5.1f}%  smp {100 * samp[k] / ts:5.1f}%  {k[0]}:{k[1]:<4d} {text[k][:84]}")
