#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.

usage: python tools/launch_summary.py gpurun_out/launches.csv [> profiles/rNN_launches_summary.txt]
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import sys


def main_timeline(path):
    """bench.py --timeline rows (name,stream,start_us,dur_us; CUPTI through torch.profiler): real concurrent durations."""
    import gzip
    agg = collections.defaultdict(lambda: [0, 0.0])
    t0, t1 = None, None
    with gzip.open(path, "rt") as f:
        next(f)
        for line in f:
            parts = line.rstrip("\n").split(",")
            name, start, dur = parts[0], float(parts[-2]), float(parts[-1])
            t0 = start if t0 is None else min(t0, start)
            t1 = start + dur if t1 is None else max(t1, start + dur)
            k = name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][-90:]
            agg[k][0] += 1
            agg[k][1] += dur
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} us summed kernel time over a {t1 - t0:.1f} us span")
    print(f"{'us':>11} {'share':>6} {'n':>6} {'avg us':>8}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:11.1f} {100 * v[1] / tot:5.1f}% {v[0]:6d} {v[1] / v[0]:8.1f}  {k}")


def main(path):
    if path.endswith(".gz"):
        return main_timeline(path)
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3, "msecond": v * 1e3, "nsecond": v / 1e3}.get(r[ui], v)
        name = r[ki].split("(")[0][-70:]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} us summed device time")
    print(f"{'us':>11} {'share':>6} {'n':>6} {'avg us':>8}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:11.1f} {100 * v[1] / tot:5.1f}% {v[0]:6d} {v[1] / v[0]:8.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
