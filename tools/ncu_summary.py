#!/usr/bin/env python
"""Print the handful of ncu metrics the profiles/ summaries quote, from an .ncu-rep (raw page as CSV).
   python tools/ncu_summary.py report.ncu-rep [extra-substring ...]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")])
        for h, u, v in zip(hdr, units, r):
            if any(h == k or h.startswith(k) for k in KEYS) or any(e in h for e in extra):
                if "stalled" in h and v and float(v.replace(",", "") or 0) < 0.05:
                    continue
                print(f"  {h} [{u}] = {v}")


if __name__ == "__main__":
    main()
