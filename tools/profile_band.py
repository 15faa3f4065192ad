#!/usr/bin/env python3
"""Render one row band of the 2048^2 dragon frame a few times (target command for ncu)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simpleraytracing_b200 import Context, Library, scene  # noqa: E402
r0, r1 = int(sys.argv[1]), int(sys.argv[2])
ppt = int(sys.argv[3]) if len(sys.argv) > 3 else 8
lib = Library(); ctx = Context(0, lib)
tri = scene.dragon()
lo, hi = lib.scene_bbox([(tri, 0.1037)])
ctx.set_meshes([(tri, 0.1037)])
ctx.set_camera(lib.camera_from_bbox(lo, hi, 2048, 2048))
L = np.zeros(2048 * 2048, np.float32)
p = lib.default_params(kernel=2, pixels_per_thread=ppt)
for _ in range(3):
    ctx.render_lbuffer(p, rows=(r0, r1), extras=False, out=dict(L=L))
    print("render_ms", ctx.timing().render_ms)
ctx.close()
