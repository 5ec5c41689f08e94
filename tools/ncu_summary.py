#!/usr/bin/env python
"""Summarise an `ncu --set full` report of one kernel launch into profiles/: a text summary (key counters, stall
reasons) and, for the wavefront kernel, the counters JSON bench.py reads (dram bytes, warp instructions).
usage: python tools/ncu_summary.py report.ncu-rep profiles/r02_wavefront [--counters] ["header line"]"""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
want_counters = "--counters" in sys.argv
header = [a for a in sys.argv[3:] if not a.startswith("--")]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units, r = rows[0], rows[1], rows[-1]
val = {k: r[i] for i, k in enumerate(h)}
unit = {k: units[i] for i, k in enumerate(h)}
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_barriers",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
lines = [f"# {x}" for x in header]
lines.append(f"# kernel: {val.get('Kernel Name', '?')}")
for k in keys:
    if k in val:
        lines.append(f"{k:78s} {val[k]} {unit.get(k, '')}")
for k in h:
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
        name = k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
        lines.append(f"  stall {name:24s} {float(val[k]):.2f}")
open(out + "_summary.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
if want_counters:
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    dram = sum(float(val[k]) * scale.get(unit[k], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    json.dump({"kernel": val.get("Kernel Name"), "dram_bytes": int(dram), "warp_instructions": int(float(val["smsp__inst_executed.sum"])),
               "duration_us_under_ncu": float(val["gpu__time_duration.sum"]), "report": rep.split("/")[-1]},
              open(out + "_counters.json", "w"), indent=1)
