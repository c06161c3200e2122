"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and share.

    python tools/summarize_launches.py gpurun_out/launches.csv "<command that was profiled>" > profiles/<name>.summary.txt
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit") == "us":
            ns *= 1e3
        elif r.get("Metric Unit") == "ms":
            ns *= 1e6
        name = re.sub(r"\(.*$", "", r["Kernel Name"])
        rows.append((name, r["Grid Size"], ns))
    tot = sum(ns for _, _, ns in rows)
    agg = collections.OrderedDict()
    for name, grid, ns in rows:
        k = (name, grid)
        c = agg.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += ns
    print(f"# {note}")
    print(f"# total {tot / 1e3:.0f} us over {len(rows)} launches (cold-cache, serialised under ncu: compare shares)")
    for (name, grid), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ns / 1e3:12.1f} us {100 * ns / tot:5.1f}%  n={n:5d} avg={ns / n / 1e3:10.2f} us  {name} grid={grid}")


if __name__ == "__main__":
    main()
