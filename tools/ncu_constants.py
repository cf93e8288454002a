#!/usr/bin/env python
"""Extract the per-frame constants bench.py quotes and the pipe metrics DESIGN.md cites from `ncu --set full` captures.

    python tools/ncu_constants.py <tag> gpurun_out/<tag>_<codec>_<kind>_<kernel>_s<streams>x<frames>.ncu-rep ...
                                                                           (run where ncu is installed, no GPU needed)

<kernel> is parameter | bank | unvoiced (the three kernels of the multi-kernel path, tools/gpu_ncu_all.sh) or fused.  For every
capture (one launch over FRAMES = streams x frames, from the file name) it writes
  profiles/<tag>_<codec>_<kind>_<kernel>_ncu_raw.txt   the raw-page metrics that matter (pipes, FP32 op counts, wavefronts,
                                                       DRAM, issue, stalls)
and merges into
  profiles/ncu_constants.json     {"<codec>/<kind>": {warp_instr_per_frame, dram_bytes_per_frame, fp32_thread_ops_per_frame
                                   (sums over the path's kernels), "kernels": {<kernel>: {warp_instr_per_frame,
                                   fp32_thread_ops_per_frame, dram_bytes_per_frame, pipe_fma_pct, pipe_alu_pct, pipe_lsu_pct,
                                   pipe_xu_pct, issue_active_pct, launch_ms_under_ncu, source}}}}
bench.py reads the sums (roofline.traffic, roofline.issue_slots).
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = re.compile(
    r"^(gpu__time_duration\.sum|launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit_\w+|waves_per_multiprocessor)"
    r"|sm__warps_active\.avg\.(pct_of_peak_sustained_active|per_cycle_active)"
    r"|smsp__inst_executed\.sum|smsp__inst_issued\.sum|sm__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active"
    r"|sm__inst_executed_pipe_(fma|fmaheavy|fmalite|alu|lsu|xu|fp64|fp16|uniform|adu|cbu|tex)\w*\.(sum|avg\.pct_of_peak_sustained_active)"
    r"|sm__pipe_(fma|fmaheavy|fmalite|alu|fp64|xu|shared)\w*_cycles_active\.avg\.pct_of_peak_sustained_active"
    r"|smsp__sass_thread_inst_executed_op_(fadd|fmul|ffma|dadd|dmul|dfma)_pred_on\.sum(\.per_cycle_elapsed)?"
    r"|smsp__sass_inst_executed_op_(shared_ld|shared_st|global_ld|global_st|local_ld|local_st)\.sum|smsp__inst_executed_op_branch\.sum"
    r"|smsp__thread_inst_executed_per_inst_executed\.ratio|sm__sass_thread_inst_executed_op_ffma_pred_on\.sum\.peak_sustained"
    r"|smsp__sass_thread_inst_executed_ops_\w+\.sum"
    r"|l1tex__data_pipe_lsu_wavefronts(_mem_shared(_op_(ld|st))?)?\.sum(\.pct_of_peak_sustained_elapsed)?"
    r"|l1tex__data_bank_conflicts_pipe_lsu_mem_shared(_op_(ld|st))?\.sum"
    r"|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed"
    r"|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__warp_issue_stalled_\w+_per_warp_active\.pct"
    r"|smsp__warps_eligible\.avg\.per_cycle_active|smsp__cycles_active\.avg|sm__cycles_elapsed\.max)$")
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def raw_page(rep):
    """rep: an .ncu-rep, or the `ncu -i rep --page raw --csv` text of one (tools/gpu_ncu_all.sh keeps only that: a report with
    imported sources is ~17 MB and gpurun brings back 64 MiB per call)"""
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    return {n: (u, v) for n, u, v in zip(names, units, vals)}


def num(m, name, default=0.0):
    if name not in m:
        return default
    u, v = m[name]
    try:
        return float(v.replace(",", "")) * UNIT.get(u, 1.0)
    except ValueError:
        return default


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    cpath = os.path.join(ROOT, "profiles", "ncu_constants.json")
    consts = json.load(open(cpath)) if os.path.exists(cpath) else {}
    for rep in reps:
        base = os.path.basename(rep)
        base = base[:-len(".ncu-rep")] if base.endswith(".ncu-rep") else base[:-len(".raw.csv")]
        m = re.match(r"(?:.*?_)?(imbe7200x4400|imbe7100x4400|ambe3600x2400|ambe3600x2450)_(hard|soft|softch|tones)_"
                     r"(parameter|bank|unvoiced|fused)_s(\d+)x(\d+)$", base)
        if not m:
            print("skip (name does not say codec_kind_kernel_s<streams>x<frames>):", rep)
            continue
        codec, kind, kname = m.group(1), m.group(2), m.group(3)
        streams, frames = int(m.group(4)), int(m.group(5))
        met = raw_page(rep)
        n = float(streams * frames)
        kernel = met.get("Kernel Name", ("", "?"))[1]
        out = os.path.join(ROOT, "profiles", "%s_%s_%s_%s_ncu_raw.txt" % (tag, codec, kind, kname))
        with open(out, "w") as f:
            f.write("# ncu --set full --clock-control none, one launch: %s\n# %s: %d streams x %d frames (%s)\n" % (kernel, base, streams, frames, rep))
            for k in sorted(met):
                if KEEP.match(k):
                    f.write("%-90s %-10s %s\n" % (k, met[k][0], met[k][1]))
        # thread-level FP32 operations (an FFMA counts two): ncu reports them per elapsed cycle, summed over the SMs
        cyc = num(met, "sm__cycles_elapsed.max")
        per_cyc = (num(met, "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed")
                   + num(met, "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed")
                   + 2.0 * num(met, "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed"))
        fp32 = per_cyc * cyc
        entry = {"warp_instr_per_frame": num(met, "smsp__inst_executed.sum") / n,
                 "fp32_thread_ops_per_frame": fp32 / n,
                 "fp32_frac_of_nonfused_issue_peak_under_ncu": per_cyc / (148.0 * 128.0),
                 "dram_bytes_per_frame": (num(met, "dram__bytes_read.sum") + num(met, "dram__bytes_write.sum")) / n,
                 "pipe_fma_pct": num(met, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", None),
                 "pipe_alu_pct": num(met, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", None),
                 "pipe_lsu_pct": num(met, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", None),
                 "pipe_xu_pct": num(met, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", None),
                 "issue_active_pct": num(met, "smsp__issue_active.avg.pct_of_peak_sustained_active", None),
                 "warps_active_pct": num(met, "sm__warps_active.avg.pct_of_peak_sustained_active", None),
                 "launch_ms_under_ncu": num(met, "gpu__time_duration.sum") if met.get("gpu__time_duration.sum", ("", ""))[0] == "ms" else None,
                 "frames_in_capture": int(n), "kernel": kernel,
                 "source": "profiles/%s_%s_%s_%s_ncu_raw.txt" % (tag, codec, kind, kname)}
        part = consts.setdefault("%s/%s" % (codec, kind), {})
        part.setdefault("kernels", {})
        if kname == "fused":
            part["fused"] = entry
        else:
            part["kernels"][kname] = entry
            ks = part["kernels"]
            for key in ("warp_instr_per_frame", "fp32_thread_ops_per_frame", "dram_bytes_per_frame"):
                part[key] = sum(ks[k][key] for k in ks)
            part["kernels_summed"] = sorted(ks)
        print("%s/%s %s: %.0f warp-instr/frame, %.0f FP32 thread-ops/frame, %.0f DRAM B/frame, fma %.1f%% alu %.1f%% lsu %.1f%% xu %.1f%% issue %.1f%%" % (
            codec, kind, kname, entry["warp_instr_per_frame"], entry["fp32_thread_ops_per_frame"], entry["dram_bytes_per_frame"],
            entry["pipe_fma_pct"] or -1, entry["pipe_alu_pct"] or -1, entry["pipe_lsu_pct"] or -1, entry["pipe_xu_pct"] or -1,
            entry["issue_active_pct"] or -1))
    with open(cpath, "w") as f:
        json.dump(consts, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
