"""Per-source-line executed-instruction / stall-sample shares of one kernel from an ncu report (built with -lineinfo).

usage: python tools/ncu_lines.py report.ncu-rep <kernel regex> [min_pct]
"""
import csv
import subprocess
import sys


def main() -> None:
    rep, regex = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{regex}",
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if "Instructions Executed" in r)
    ie, samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    items, tot_i, tot_s = [], 0, 0
    for r in rows:
        if r and r[0].isdigit() and len(r) > ie:
            try:
                n, s = int(r[ie]), int(r[samp])
            except ValueError:
                continue
            items.append((int(r[0]), n, s, r[1].strip()))
            tot_i += n
            tot_s += s
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    for line, n, s, text in items:
        if 100 * n / max(tot_i, 1) >= min_pct or 100 * s / max(tot_s, 1) >= min_pct:
            print(f"{100 * n / tot_i:5.1f}% inst {100 * s / tot_s:5.1f}% stall  L{line:<4d} {text[:120]}")


if __name__ == "__main__":
    main()
