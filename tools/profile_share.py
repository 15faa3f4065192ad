#!/usr/bin/env python3
"""Target of ncu launch lists: the kernels of one block-cyclic share of a frame (what a rank renders at N GPUs).
usage: profile_share.py [n_parts] [kernel selector] [reps] [size] [scan residency cap]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simpleraytracing_b200 import Context, Library, scene  # noqa: E402
n_parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
kernel = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
size = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
cap = int(sys.argv[5]) if len(sys.argv) > 5 else 0
lib = Library(); ctx = Context(0, lib)
tri = scene.dragon()
lo, hi = lib.scene_bbox([(tri, 0.1037)])
ctx.set_meshes([(tri, 0.1037)]); ctx.set_camera(lib.camera_from_bbox(lo, hi, size, size)); ctx.set_scan_residency(cap)
L = torch.zeros(size * size, dtype=torch.float32, device="cuda"); img = torch.zeros_like(L)
p = lib.default_params(kernel=kernel)
for _ in range(reps):
    if n_parts == 1:
        ctx.render_lbuffer_device(p, (0, size), L.data_ptr())
    else:
        ctx.render_lbuffer_blocks_device(p, n_parts, 1, [L.data_ptr()])
    ctx.repair_device(p, L.data_ptr(), img.data_ptr())
    t = ctx.timing()
    print(f"share 1/{n_parts} kernel {t.kernel_used}: pass {t.render_ms:.4f} ms, main kernel {t.main_kernel_ms:.4f} ms, repair {t.repair_ms:.4f}", flush=True)
ctx.close()
