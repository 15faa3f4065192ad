#!/usr/bin/env python3
"""Time the filtered kernel of one library build: full 2048^2 frame + top / centre 256-row bands (development aid)."""
import json, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simpleraytracing_b200 import Context, Library, scene, KERNEL_FILTERED  # noqa: E402
lib = Library(os.environ.get("SRT_LIB") or None); ctx = Context(0, lib)
tri = scene.dragon(); nt = len(tri)
lo, hi = lib.scene_bbox([(tri, 0.1037)])
ctx.set_meshes([(tri, 0.1037)])
size = int(os.environ.get("SRT_SIZE", "2048"))
ctx.set_camera(lib.camera_from_bbox(lo, hi, size, size))
L = np.zeros(size * size, np.float32)
out = {"lib": os.environ.get("SRT_LIB", "default")}
for ppt in [int(v) for v in os.environ.get("SRT_PPTS", "8").split(",")]:
    p = lib.default_params(kernel=KERNEL_FILTERED, pixels_per_thread=ppt)
    for name, rows in (("full", (0, size)), ("top256", (0, 256)), ("mid256", ((size - 256) // 2, (size + 256) // 2))):
        ms = []
        for _ in range(4):
            ctx.render_lbuffer(p, rows=rows, extras=False, out=dict(L=L))
            ms.append(ctx.timing().render_ms)
        out[f"P{ppt}_{name}"] = round(min(ms[1:]), 3)
print(json.dumps(out), flush=True)
ctx.close()
