#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full capture) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof_nn.ncu-rep [kernel-substring] > profiles/ncu_nn_rNN.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
want_kernel = sys.argv[2] if len(sys.argv) > 2 else ""
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print(f"# ncu --set full --clock-control none summary of {rep} (raw page); one block per captured launch")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if want_kernel not in name:
        continue
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:70s} {r[i][:150]} {units[i]}")
    stalls = [(hdr[i], float(r[i])) for i in range(len(hdr)) if hdr[i].startswith("smsp__average_warp") and "issue_stalled" in hdr[i] and r[i] not in ("", "n/a")]
    stalls = [(h, v) for h, v in stalls if "_not_issued" not in h]
    for h, v in sorted(stalls, key=lambda t: -t[1])[:8]:
        print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', '')[:60]:62s} {v:.3f}")
    print("-" * 100)
