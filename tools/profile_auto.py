#!/usr/bin/env python3
"""Render the 2048^2 dragon frame with the default (auto) kernel selection a few times (target command for ncu)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simpleraytracing_b200 import Context, Library, scene  # noqa: E402
size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
lib = Library(); ctx = Context(0, lib)
tri = scene.dragon(); lo, hi = lib.scene_bbox([(tri, 0.1037)])
ctx.set_meshes([(tri, 0.1037)]); ctx.set_camera(lib.camera_from_bbox(lo, hi, size, size))
buf = torch.zeros(size * size, dtype=torch.float32, device="cuda")
img = torch.zeros(size * size, dtype=torch.float32, device="cuda")
p = lib.default_params()
for _ in range(3):
    ctx.render_lbuffer_device(p, (0, size), buf.data_ptr())
    ctx.repair_device(p, buf.data_ptr(), img.data_ptr())
    print("render_ms", ctx.timing().render_ms, flush=True)
ctx.close()
