#!/usr/bin/env python3
"""Full 2048^2 frame and contiguous bands with forced triangle-range splits (development aid)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simpleraytracing_b200 import Context, Library, scene  # noqa: E402
lib = Library(); ctx = Context(0, lib)
tri = scene.dragon(); lo, hi = lib.scene_bbox([(tri, 0.1037)])
ctx.set_meshes([(tri, 0.1037)]); ctx.set_camera(lib.camera_from_bbox(lo, hi, 2048, 2048))
buf = torch.zeros(2048 * 2048, dtype=torch.float32, device="cuda")
p = lib.default_params(kernel=2)
for rows in ((0, 2048), (896, 1152), (0, 256)):
    for zs in (0, 6, 12, 24):
        ctx.force_splits(zs)
        ms = []
        for _ in range(4):
            ctx.render_lbuffer_device(p, rows, buf.data_ptr()); ms.append(ctx.timing().render_ms)
        print("rows", rows, "splits", zs, "ms", round(min(ms[1:]), 4), flush=True)
ctx.force_splits(0)
ctx.close()
