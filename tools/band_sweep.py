#!/usr/bin/env python3
"""Render-kernel time vs band height and pixels-per-thread (development aid for the multi-GPU grid sizing)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simpleraytracing_b200 import Context, Library, scene, KERNEL_FILTERED  # noqa: E402

lib = Library(); ctx = Context(0, lib)
tri = scene.dragon(); nt = len(tri)
lo, hi = lib.scene_bbox([(tri, 0.1037)])
ctx.set_meshes([(tri, 0.1037)])
size = 2048
ctx.set_camera(lib.camera_from_bbox(lo, hi, size, size))
L = np.zeros(size * size, np.float32)
for rows in (128, 256, 512, 1024, 2048):
    for start in sorted({0, (size - rows) // 2}):
        for ppt in range(1, 9):
            p = lib.default_params(kernel=KERNEL_FILTERED, pixels_per_thread=ppt)
            ms = []
            for it in range(4):
                ctx.render_lbuffer(p, rows=(start, start + rows), extras=False, out=dict(L=L))
                ms.append(ctx.timing().render_ms)
            best = min(ms[1:])
            print(json.dumps({"rows": rows, "start": start, "ppt": ppt, "ms": round(best, 4),
                              "ms_per_256rows": round(best * 256 / rows, 4),
                              "tests_per_s": rows * size * nt / (best * 1e-3)}), flush=True)
ctx.close()
