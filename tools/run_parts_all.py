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
p = lib.default_params(kernel=2)
out = []
for part in range(n_parts):
    ms = []
    for _ in range(4):
        ctx.render_lbuffer_interleaved_device(p, n_parts, part, slot.data_ptr()); ms.append(ctx.timing().render_ms)
    out.append(round(min(ms[1:]), 4))
print("parts", n_parts, "per-part ms", out, "sum", round(sum(out), 3), flush=True)
for zs in (4, 8, 12, 16, 24, 32, 48):
    ctx.force_splits(zs)
    ms = []
    for _ in range(4):
        ctx.render_lbuffer_interleaved_device(p, n_parts, 0, slot.data_ptr()); ms.append(ctx.timing().render_ms)
    print("part 0 splits", zs, "ms", round(min(ms[1:]), 4), flush=True)
ctx.force_splits(0)
ctx.close()
