"""Exports the raw page of an ncu report to a compact text summary (one block per launch, selected metrics).

usage: python tools/ncu_export.py report.ncu-rep > summary.txt   (runs `ncu -i`, so it works on the GPU box and here)
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
]


def main() -> None:
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"{r[idx['Kernel Name']][:90]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        for w in WANT:
            if w in idx:
                print(f"    {w} = {r[idx[w]]} {units[idx[w]]}")


if __name__ == "__main__":
    main()
