#!/usr/bin/env python
"""Write a text summary of one ncu --set full capture: key raw metrics, stall reasons and (for k_run) the per-function table.
usage: tools/ncu_summary.py <report.ncu-rep> <out.txt> <title> [<lib.so> <kernel>]"""
import csv, io, json, subprocess, sys

rep, out, title = sys.argv[1:4]
o = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(o)))
H, U = rows[0], rows[1]
lines = ["# " + title]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
for k, V in enumerate(rows[2:]):
    d = {h: (v, u) for h, u, v in zip(H, U, V)}
    lines.append(f"## launch {k}: {d.get('Kernel Name', ('?', ''))[0][:90]}")
    for w in want:
        if w in d:
            lines.append(f"{w:72s} {d[w][0]} {d[w][1]}")
    st = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v) for h, u, v in zip(H, U, V)
          if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")}
    lines.append("stall cycles per issued instruction: " + ", ".join(f"{k}={v:.2f}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:7]))
if len(sys.argv) > 5:
    t = subprocess.run([sys.executable, "tools/ncu_by_function.py", rep, sys.argv[4], sys.argv[5]], capture_output=True, text=True).stdout
    lines.append("## per-function breakdown (tools/ncu_by_function.py, same build)")
    lines += t.splitlines()[:28]
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
