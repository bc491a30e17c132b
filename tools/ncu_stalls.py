"""Warp-stall reasons of one kernel from an ncu report: totals per reason and the instructions that collect the most samples.

usage: python tools/ncu_stalls.py report.ncu-rep <kernel regex> [top_n]
"""
import csv
import subprocess
import sys


def main() -> None:
    rep, regex = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--kernel-name", f"regex:{regex}",
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if "# Samples" in r)
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot, n, items = {s: 0 for s in stalls}, 0, []
    for r in rows:
        if len(r) < len(hdr) or not r[idx["# Samples"]].isdigit():
            continue
        smp = int(r[idx["# Samples"]])
        n += smp
        d = {s: int(r[idx[s]]) for s in stalls if r[idx[s]].isdigit() and int(r[idx[s]]) > 0}
        for s, v in d.items():
            tot[s] += v
        items.append((smp, r[idx["Source"]].strip(), d))
    print(rows[0][1][:100] if rows and len(rows[0]) > 1 else "")
    print(f"total samples {n}")
    for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  {s:24s} {100 * v / max(n, 1):5.1f}%")
    for smp, txt, d in sorted(items, key=lambda t: -t[0])[:top]:
        why = ", ".join(f"{k[6:]} {v}" for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:3])
        print(f"{100 * smp / max(n, 1):5.1f}%  {txt[:64]:64s} {why}")


if __name__ == "__main__":
    main()
