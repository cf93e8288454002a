#!/usr/bin/env python
"""Key metrics of one kernel launch from an .ncu-rep (ncu --page raw --csv): pipes, issue, occupancy, stalls, memory.
usage: python tools/ncu_key_metrics.py report.ncu-rep [frames_in_launch]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
v = rows[2] if len(rows) > 2 else rows[1]
m = dict(zip(h, v))
print("kernel:", m.get("Kernel Name"), " grid", m.get("launch__grid_size"), "block", m.get("launch__block_size"),
      "regs", m.get("launch__registers_per_thread"), "dyn smem", m.get("launch__shared_mem_per_block_dynamic"))
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__maximum_warps_per_active_cycle_pct"]
for k in keys:
    if k in m:
        print("%-78s %s" % (k, m[k]))
if frames and "smsp__inst_executed.sum" in m:
    print("warp-instructions per frame: %.0f" % (float(m["smsp__inst_executed.sum"].replace(",", "")) / frames))
st = [(float(m[k]), k) for k in m if "average_warps_issue_stalled" in k and k.endswith("_per_issue_active.ratio") and "not_issued" not in k]
tot = sum(x for x, _ in st)
print("stall reasons (warps per issue slot; share):")
for x, k in sorted(st, reverse=True)[:9]:
    print("   %-28s %6.2f  %4.1f %%" % (k.split("issue_stalled_")[1].split("_per_issue")[0], x, 100 * x / tot))
