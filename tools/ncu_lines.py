"""Per-source-line stall samples and executed instructions from `ncu --page source --csv
--print-source cuda,sass` output.  usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
iline, isrc = 0, 1
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
samp, ex, text = defaultdict(int), defaultdict(int), {}
line = None
for r in rows[hi + 1:]:
    if len(r) <= isamp:
        continue
    if r[iline].strip().isdigit():
        line = int(r[iline])
        text[line] = r[isrc]
    if line is None:
        continue
    if r[isamp].strip().isdigit():
        samp[line] += int(r[isamp])
    if r[iex].strip().isdigit():
        ex[line] += int(r[iex])
tot = sum(samp.values()) or 1
tex = sum(ex.values()) or 1
print(f"total samples {tot}, warp instructions {tex}")
for l in sorted(samp, key=lambda l: -samp[l])[:top]:
    print(f"{samp[l]:7d} {100 * samp[l] / tot:5.1f}%  instr {100 * ex[l] / tex:5.1f}%  L{l:<5d} {text.get(l, '')[:100]}")
