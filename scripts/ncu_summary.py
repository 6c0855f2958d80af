"""Compact metric,unit,value summary of one kernel from an .ncu-rep (ncu --set full capture).
Usage: python scripts/ncu_summary.py capture.ncu-rep > profiles/<name>_ncu_summary.csv"""
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tc", "sm__pipe_tc", "sm__inst_executed_pipe_uniform",
        "sm__inst_executed_pipe_tmem", "smsp__inst_executed_pipe_tmem", "sm__pipe_tensor", "sm__inst_executed_pipe_tensor")
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
print("metric,unit,value")
print(f"kernel,,\"{vals[hdr.index('Kernel Name')]}\"")
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(k) for k in KEEP) and v != "" and "TriageCompute" not in h:
        print(f"{h},{u},{v}")
