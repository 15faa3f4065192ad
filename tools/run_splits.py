#!/usr/bin/env python3
"""Time one block-cyclic share of the 2048^2 dragon frame with forced triangle-range splits (development aid)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simpleraytracing_b200 import Context, Library, scene  # noqa: E402
lib = Library(); ctx = Context(0, lib)
tri = scene.dragon(); lo, hi = lib.scene_bbox([(tri, 0.1037)])
ctx.set_meshes([(tri, 0.1037)]); ctx.set_camera(lib.camera_from_bbox(lo, hi, 2048, 2048))
for n_parts in (8, 4, 2):
    sr = lib.interleave_slot_rows(2048, n_parts)
    slot = torch.zeros(sr * 2048, dtype=torch.float32, device="cuda")
    for zs in [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,1,2,3,4".split(","))]:
        ctx.force_splits(zs)
        p = lib.default_params(kernel=2, pixels_per_thread=0 if zs == 0 else 0)
        ms = []
        for _ in range(4):
            ctx.render_lbuffer_interleaved_device(p, n_parts, 0, slot.data_ptr()); ms.append(ctx.timing().render_ms)
        print("parts", n_parts, "splits", zs, "ms", round(min(ms[1:]), 4), "launches", ctx.timing().launches, flush=True)
ctx.force_splits(0)
ctx.close()
