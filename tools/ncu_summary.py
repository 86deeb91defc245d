#!/usr/bin/env python
"""Condenses an `ncu --set full` report into the per-launch table kept under profiles/:

    python tools/ncu_summary.py gpurun_out/r02_prof.ncu-rep profiles/r02_ncu_full_summary.csv

One row per captured launch: duration, DRAM bytes, DRAM / L2 / SM throughput, tensor-pipe activity, occupancy limits,
executed instructions and issue-slot use, L2 -> SM bytes.  (`ncu -i <rep> --page raw --csv` is the source.)"""
import csv
import subprocess
import sys

COLUMNS = [
    "Kernel Name", "Grid Size", "Block Size",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units = rows[0], rows[1]
    idx = [head.index(c) for c in COLUMNS if c in head]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([head[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            if len(r) == len(head):
                w.writerow([r[i] for i in idx])
    print(out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main()
