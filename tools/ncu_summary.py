"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel and per (kernel, grid) shape.

usage: python tools/ncu_summary.py gpurun_out/launches.csv [--by-grid] > profiles/rN_ncu_launch_summary.txt
"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name.replace("rf::", "")[:96]


def main() -> None:
    path = sys.argv[1]
    by_grid = "--by-grid" in sys.argv
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit") in ("us", "usecond"):
            ns *= 1e3
        rows.append((short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], ns))
    total = sum(r[3] for r in rows)
    print(f"total {total / 1e6:.2f} ms over {len(rows)} launches")
    agg = defaultdict(lambda: [0.0, 0])
    for name, grid, block, ns in rows:
        key = (name, grid) if by_grid else (name,)
        agg[key][0] += ns
        agg[key][1] += 1
    for key, (ns, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        label = key[0] + (f"  grid {key[1]}" if by_grid else "")
        print(f"  {ns / 1e6:8.3f} ms  {100 * ns / total:5.1f}%  {n:5d} launches  {ns / n / 1e3:8.1f} us/launch  {label}")


if __name__ == "__main__":
    main()
