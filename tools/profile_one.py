#!/usr/bin/env python3
"""Run the filtered render kernel a few times at one size (target command for ncu)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simpleraytracing_b200 import Context, Library, scene  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
kernel = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ppt = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
lib = Library(); ctx = Context(0, lib)
tri = scene.dragon()
lo, hi = lib.scene_bbox([(tri, 0.1037)])
ctx.set_meshes([(tri, 0.1037)])
ctx.set_camera(lib.camera_from_bbox(lo, hi, size, size))
L = np.zeros(size * size, np.float32)
p = lib.default_params(kernel=kernel, pixels_per_thread=ppt)
for _ in range(reps):
    ctx.render_lbuffer(p, extras=False, out=dict(L=L))
    print("render_ms", ctx.timing().render_ms)
ctx.close()
