"""Aggregate an ncu --csv launch list (gpu__time_duration.sum) per kernel: launches, total, average, share."""
import collections
import csv
import re
import sys


def main(path):
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in csv.reader(open(path, errors="ignore")):
        if "Kernel Name" in r:
            hdr = r
            continue
        if not hdr or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(d["Metric Unit"], 1.0)
        k = re.sub(r"\(.*", "", d["Kernel Name"])
        a = agg[k]
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    tot = sum(v[1] for v in agg.values()) or 1.0
    print(f"{'kernel':44s} {'n':>5s} {'total us':>10s} {'avg us':>8s} {'max us':>8s} share")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:44s} {v[0]:5d} {v[1]:10.1f} {v[1] / v[0]:8.1f} {v[2]:8.1f} {100 * v[1] / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
