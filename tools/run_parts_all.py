#!/usr/bin/env python3
"""Kernel time of every block-cyclic share of the 2048^2 dragon frame on one GPU (load balance across ranks) and of
share 0 with forced triangle-range splits (development aid)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simpleraytracing_b200 import Context, Library, scene  # noqa: E402
n_parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lib = Library(); ctx = Context(0, lib)
tri = scene.dragon(); lo, hi = lib.scene_bbox([(tri, 0.1037)])
ctx.set_meshes([(tri, 0.1037)]); ctx.set_camera(lib.camera_from_bbox(lo, hi, 2048, 2048))
sr = lib.interleave_slot_rows(2048, n_parts)
slot = torch.zeros(sr * 2048, dtype=torch.float32, device="cuda")
ppt = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if ppt:
    ctx.force_splits(-1)            # pixels_per_thread picks the run length of the split launches
p = lib.default_params(kernel=2, pixels_per_thread=ppt)
out = []
for part in range(n_parts):
    ms = []
    for _ in range(4):
        ctx.render_lbuffer_interleaved_device(p, n_parts, part, slot.data_ptr()); ms.append(ctx.timing().render_ms)
    out.append(round(min(ms[1:]), 4))
print("parts", n_parts, "ppt", ppt, "per-part ms", out, "sum", round(sum(out), 3), "max", max(out), flush=True)
for zs in (12, 16, 24, 32):
    ctx.force_splits(zs)
    ms = []
    for _ in range(4):
        ctx.render_lbuffer_interleaved_device(p, n_parts, 0, slot.data_ptr()); ms.append(ctx.timing().render_ms)
    print("part 0 splits", zs, "ms", round(min(ms[1:]), 4), flush=True)
ctx.force_splits(0)
ctx.close()
