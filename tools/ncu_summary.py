#!/usr/bin/env python3
"""Writes a compact, committable text summary of an .ncu-rep (`ncu --set full` capture):
per kernel the metrics the roofline / issue analysis uses, then the top CUDA source lines by
executed warp instructions.   usage: ncu_summary.py report.ncu-rep out.txt [top_n]"""
import csv
import subprocess
import sys
from pathlib import Path

rep, out = sys.argv[1], sys.argv[2]
top = sys.argv[3] if len(sys.argv) > 3 else "25"
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
lines = [f"# {Path(rep).name}: ncu --set full --clock-control none (cold-cache, serialised replays)"]
names = []
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1]
    names.append(name)
    lines.append(f"\n== {name}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"  {k:82s} {r[i]} {units[i]}")
here = Path(__file__).resolve().parent
for name in dict.fromkeys(names):
    t = subprocess.run([sys.executable, str(here / "ncu_lines.py"), rep, name, top], capture_output=True, text=True).stdout
    lines.append(f"\n-- top source lines of {name} (share of warp instructions, share of stall samples, avg active threads)")
    lines.append(t.rstrip())
Path(out).write_text("\n".join(lines) + "\n")
print(f"wrote {out}: {len(lines)} lines")
