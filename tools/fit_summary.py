"""Prints the fit part of a bench.py JSON line (development helper)."""
import json
import sys

line = json.load(open(sys.argv[1]))
f = line["fit"]
print({k: round(v, 3) for k, v in f.items() if isinstance(v, float)})
print(f["phases_ms"])
print("K1", line["ms_per_step"], line["roofline"]["frac"])
