"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
Usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/launches_rNN.md"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}.get(unit, 1.0)
        rows.append((name, ns))
    tot = sum(ns for _, ns in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for name, ns in rows:
        short = re.sub(r"\(.*$", "", name)
        short = re.sub(r"^void ", "", short)
        short = re.sub(r"\(anonymous namespace\)::", "", short)
        agg[short][0] += 1
        agg[short][1] += ns
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% | %.1f |" % (name[:90], n, ns / 1e6, 100 * ns / tot, ns / n / 1e3))
    print("\ntotal: %d launches, %.3f ms (cold-cache, serialised: compare SHARES, not absolutes)" % (len(rows), tot / 1e6))


if __name__ == "__main__":
    main(sys.argv[1])
