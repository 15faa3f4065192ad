#!/usr/bin/env python3
"""Time one frame size with each kernel selector (device time from the library's events). Development aid."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simpleraytracing_b200 import Context, Library, scene, KERNEL_FILTERED, KERNEL_CULLED  # noqa: E402
size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
which = sys.argv[2] if len(sys.argv) > 2 else "dragon"
kernels = [int(k) for k in sys.argv[3].split(",")] if len(sys.argv) > 3 else [KERNEL_CULLED, KERNEL_FILTERED]
lib = Library(); ctx = Context(0, lib)
tri = scene.dragon() if which == "dragon" else scene.uv_sphere()
lo, hi = lib.scene_bbox([(tri, 0.1037)])
ctx.set_meshes([(tri, 0.1037)]); ctx.set_camera(lib.camera_from_bbox(lo, hi, size, size))
buf = torch.zeros(size * size, dtype=torch.float32, device="cuda")
ctx.count_exact(True) if hasattr(ctx, "count_exact") else None
for k in kernels:
    p = lib.default_params(kernel=k)
    ts = []
    for _ in range(5):
        ctx.render_lbuffer_device(p, (0, size), buf.data_ptr())
        t = ctx.timing()
        ts.append(t.render_ms)
    print(f"kernel {k} {which} {size}^2: render_ms {['%.3f' % v for v in ts]} exact_tests {t.exact_tests} pairs {t.pair_tests}", flush=True)
ctx.close()
