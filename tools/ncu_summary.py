#!/usr/bin/env python
"""Markdown summary of an .ncu-rep (one kernel): the counters DESIGN.md / bench.py quote and the
hot SASS blocks.  Usage: ncu_summary.py report.ncu-rep "title" > profiles/xxx.md"""
import csv, subprocess, sys, io, collections
rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
print("# %s\n" % title)
print("`%s`\n" % name)
print("| metric | value |\n|---|---|")
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("| %s | %s %s |" % (w, vals[i], units[i]))
out = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_sass.py"), rep, "14"],
                     capture_output=True, text=True).stdout
print("\n## SASS: executed instructions by opcode and the hottest blocks\n\n```\n" + out + "```")
