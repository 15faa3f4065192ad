#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full --import-source on) into a text file for profiles/.
usage: tools/ncu_summary.py report.ncu-rep out.txt [launch index, default 0]  (the source section covers launch 0 only)"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # which captured launch (raw page row) to summarise
vals = rows[2 + which]
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_issued.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
keys += [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
with open(out, "w") as f:
    f.write(f"# ncu summary of {rep}\n")
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            f.write(f"{k:90s} {vals[i]} {units[i]}\n")
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 2 and which == 0:
        h = rows[1]
        body = []
        for r in rows[2:]:                      # first kernel only (a second 'Kernel Name' header starts the next one)
            if len(r) != len(h):
                break
            body.append(r)
        ix = {c: i for i, c in enumerate(h)}
        cols = ["stall_dispatch", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_math", "stall_barrier",
                "stall_branch_resolving", "stall_no_inst", "stall_not_selected", "stall_selected"]
        mx = max(int(r[ix["Instructions Executed"]]) for r in body)
        f.write(f"\n# hot loop (instructions executed >= 98% of the max = {mx} warp-level executions), "
                "warp-stall samples per SASS instruction\n")
        for r in body:
            if int(r[ix["Instructions Executed"]]) >= 0.98 * mx:
                f.write(f"{r[1].strip()[:56]:58s} samples={int(r[ix['# Samples']]):7d} " +
                        " ".join(f"{c[6:]}={r[ix[c]]}" for c in cols if r[ix[c]] != "0") + "\n")
print("wrote", out)
