#!/usr/bin/env python3
"""Render one block-cyclic part of the 2048^2 dragon frame a few times (target command for ncu / timing)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simpleraytracing_b200 import Context, Library, scene  # noqa: E402
n_parts, part = int(sys.argv[1]), int(sys.argv[2])
ppt = int(sys.argv[3]) if len(sys.argv) > 3 else -1
size = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
lib = Library(); ctx = Context(0, lib)
tri = scene.dragon()
lo, hi = lib.scene_bbox([(tri, 0.1037)])
ctx.set_meshes([(tri, 0.1037)])
ctx.set_camera(lib.camera_from_bbox(lo, hi, size, size))
sr = lib.interleave_slot_rows(size, n_parts)
slot = torch.zeros(sr * size, dtype=torch.float32, device="cuda")
for ppt_i in ([ppt] if ppt >= 0 else [0, 8, 7, 6, 5, 4]):
    p = lib.default_params(kernel=2, pixels_per_thread=ppt_i)
    ms = []
    for _ in range(4):
        ctx.render_lbuffer_interleaved_device(p, n_parts, part, slot.data_ptr())
        ms.append(ctx.timing().render_ms)
    print("ppt", ppt_i, "render_ms", round(min(ms[1:]), 4), flush=True)
ctx.close()
