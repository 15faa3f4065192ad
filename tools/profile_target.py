#!/usr/bin/env python3
"""Target command of the ncu captures: a few frames of one workload through one kernel selector.
usage: profile_target.py [c2|c3|c4|c5] [kernel selector: 2 filtered, 4 culled] [reps]"""
import sys
from pathlib import Path
import argparse
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from simpleraytracing_b200 import Context, Library  # noqa: E402
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
kernel = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
wl = bench.workload(argparse.Namespace(config=cfg, size=0))
lib = Library(); ctx = Context(0, lib)
lo, hi = lib.scene_bbox([(wl["cam_tri"], 0.0)])
ctx.set_meshes(wl["meshes"]); ctx.set_camera(lib.camera_from_bbox(lo, hi, wl["W"], wl["H"]))
W, H = wl["W"], wl["H"]
L = torch.zeros(W * H, dtype=torch.float32, device="cuda"); img = torch.zeros_like(L)
p = lib.default_params(kernel=kernel)
for _ in range(reps):
    ctx.render_lbuffer_device(p, (0, H), L.data_ptr()); ctx.repair_device(p, L.data_ptr(), img.data_ptr())
    t = ctx.timing()
    print(f"{cfg} kernel {t.kernel_used}: pass {t.render_ms:.4f} ms, main kernel {t.main_kernel_ms:.4f} ms, scan_cols {t.scan_cols}", flush=True)
ctx.close()
