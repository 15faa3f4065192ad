#!/usr/bin/env python3
"""Extract the per-kernel counters bench.py's roofline block quotes from an `ncu --set full` report.

  tools/ncu_counters.py report.ncu-rep profiles/r02_kernel_counters.json --width 2048 --triangles 22866 [--n-gpus 1] [--rows R]

For every captured launch: warp instructions executed (smsp__inst_executed.sum), DRAM bytes read + written, duration,
issue-slot utilisation, pipe utilisations.  `warp_inst_per_test` = instructions / (rows x width x triangles): the
(ray, triangle) tests the launch resolves (rows = the image rows of the launch; default: the whole frame)."""
import argparse
import csv
import io
import json
import subprocess
from pathlib import Path

ap = argparse.ArgumentParser()
ap.add_argument("report"); ap.add_argument("out")
ap.add_argument("--width", type=int, required=True); ap.add_argument("--triangles", type=int, required=True)
ap.add_argument("--rows", type=int, default=0); ap.add_argument("--n-gpus", type=int, default=1)
ap.add_argument("--box", default=""); ap.add_argument("--source", default="")
args = ap.parse_args()
raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}


def num(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return None


out = Path(args.out)
doc = json.loads(out.read_text()) if out.exists() else {"captures": []}
tests = float(args.rows or args.width) * args.width * args.triangles
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[ix["Kernel Name"]].split("(")[0].split("<")[0].split("::")[-1].strip()
    inst = num(r, "smsp__inst_executed.sum")
    unit_r = rows[1][ix["dram__bytes_read.sum"]]; unit_w = rows[1][ix["dram__bytes_write.sum"]]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dr = (num(r, "dram__bytes_read.sum") or 0.0) * scale.get(unit_r, 1.0)
    dw = (num(r, "dram__bytes_write.sum") or 0.0) * scale.get(unit_w, 1.0)
    e = {"kernel": name, "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]], "width": args.width, "n_gpus": args.n_gpus,
         "rows": args.rows or args.width, "triangles": args.triangles,
         "box": [int(v) for v in args.box.split("x")] if args.box else None,
         "duration_us": (num(r, "gpu__time_duration.sum") or 0.0) / (1e3 if rows[1][ix["gpu__time_duration.sum"]] in ("ns", "nsecond") else 1.0),
         "warp_inst": inst, "warp_inst_per_test": inst / tests if inst else None,
         "dram_bytes": dr + dw, "dram_bytes_read": dr, "dram_bytes_write": dw,
         "issue_slots_pct": num(r, "sm__inst_issued.avg.pct_of_peak_sustained_active") or num(r, "sm__instruction_throughput.avg.pct_of_peak_sustained_active"),
         "ipc_per_sm": num(r, "sm__inst_executed.avg.per_cycle_elapsed"),
         "fma_pipe_pct": num(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
         "alu_pipe_pct": num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
         "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
         "registers": num(r, "launch__registers_per_thread"),
         "l1_hit_pct": num(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": num(r, "lts__t_sector_hit_rate.pct"),
         "source": args.source or f"ncu --set full capture {Path(args.report).name}"}
    doc["captures"] = [c for c in doc["captures"] if not (c["kernel"] == name and c["width"] == args.width and c["n_gpus"] == args.n_gpus and c["rows"] == e["rows"])]
    doc["captures"].append(e)
out.write_text(json.dumps(doc, indent=1) + "\n")
print("wrote", out, [c["kernel"] for c in doc["captures"]])
