#!/usr/bin/env python3
"""Full 2048^2 frame with forced triangle-range splits: isolates the SPLIT kernel's per-test cost (development aid)."""
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
for zs in (0, 1, 2, 3, 4):
    ctx.force_splits(zs)
    ms = []
    for _ in range(4):
        ctx.render_lbuffer_device(p, (0, 2048), buf.data_ptr()); ms.append(ctx.timing().render_ms)
    print("full frame splits", zs, "ms", round(min(ms[1:]), 3), flush=True)
ctx.force_splits(0); ctx.close()
