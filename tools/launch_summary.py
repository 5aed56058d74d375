#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launch count per kernel.

    python tools/launch_summary.py gpurun_out/launches.csv "title" > profiles/rNN_launches.md
Per-launch times under ncu are cold-cache and serialised: the SHARE of a kernel is what is comparable with bench.py."""
import csv
import sys


def main():
    path, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = {}
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else (v if unit.startswith("m") else v * 1e3))
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0, row["Grid Size"], row["Block Size"]])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    print(f"# {title}\n\nSource: `{path}` ({sum(a[0] for a in agg.values())} launches, {total:.3f} ms of kernel time under ncu).\n")
    print("| kernel | launches | total ms | share | last grid | block |\n|---|---|---|---|---|---|")
    for name, (cnt, ms, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {cnt} | {ms:.3f} | {100 * ms / total:.1f} % | {grid} | {block} |")


if __name__ == "__main__":
    main()
