"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the forward)."""
import collections
import csv
import sys


def load(path):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(lines[start:]))


def short(name):
    base = name.split("(")[0].split("::")[-1]
    if "<" in name and ("gemm_tc" in name or "attn_tc" in name or "conv" in name):
        base = name.split("(")[0].split("::")[-1]
    return base


if __name__ == "__main__":
    rows = load(sys.argv[1])
    tot, by = 0.0, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        n = short(r["Kernel Name"])
        t = float(r["Metric Value"]) / 1e3
        by[n][0] += 1
        by[n][1] += t
        tot += t
    for n, (c, t) in sorted(by.items(), key=lambda x: -x[1][1]):
        print(f"{n:60s} {c:4d} launches {t:9.1f} us {100 * t / tot:5.1f}%  avg {t / c:7.1f} us")
    print(f"total {tot:.1f} us over {len(rows)} launches")
    if len(sys.argv) > 2:  # per-launch dump
        for r in rows:
            print(r["ID"], short(r["Kernel Name"]), r["Grid Size"], r["Metric Value"])
