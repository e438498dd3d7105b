#!/usr/bin/env python
"""ncu --metrics gpu__time_duration.sum --csv launch list -> per-kernel totals and shares."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3}.get(row["Metric Unit"], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.3f} ms total (cold-cache, serialised)")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ms:12.3f} ms {100 * ms / tot:6.2f}%  n={n:4d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
