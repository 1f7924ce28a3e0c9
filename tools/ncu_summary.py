#!/usr/bin/env python
"""Condense an ncu report (--set full) into the handful of numbers DESIGN.md / bench.py quote.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sector_hit_rate.pct",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of {rep} (ncu --set full --clock-control none; cold-cache, serialised replays)")
    for r in rows[2:]:
        print(f"\nkernel: {r[hdr.index('Kernel Name')]}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        for k in KEYS:
            if k in hdr:
                print(f"  {k:70s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
        stalls = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr) if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i]]
        for v, h in sorted(stalls, reverse=True)[:6]:
            print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:.3f} warps/issue")


if __name__ == "__main__":
    main()
