"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path, width=90):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        a = agg.setdefault(row["Kernel Name"][:width], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t:12.1f} us {n:6d} launches {100 * t / tot:5.1f}%  {k}")
    print(f"{tot:12.1f} us total")


if __name__ == "__main__":
    main(sys.argv[1])
