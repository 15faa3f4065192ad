#!/usr/bin/env python3
"""Kernel timing sweep on one GPU (development aid): sizes x pixels-per-thread x kernel."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simpleraytracing_b200 import Context, Library, scene, KERNEL_EXACT, KERNEL_FILTERED  # noqa: E402


def main():
    import os
    lib = Library(os.environ.get("SRT_LIB") or None); ctx = Context(0, lib)
    only_filtered = bool(os.environ.get("SRT_ONLY_FILTERED"))
    ppt_list = tuple(int(v) for v in os.environ.get("SRT_PPTS", "1,2,4,8").split(","))
    tri = scene.dragon(); nt = len(tri)
    lo, hi = lib.scene_bbox([(tri, 0.1037)])
    ctx.set_meshes([(tri, 0.1037)])
    print(json.dumps({"ffma_peak_tflops": ctx.ffma_peak_tflops(3)}), flush=True)
    sizes = [int(s) for s in sys.argv[1:]] or [1024, 2048]
    for size in sizes:
        ctx.set_camera(lib.camera_from_bbox(lo, hi, size, size))
        L = np.zeros(size * size, np.float32)
        for kernel, ppts in ((KERNEL_FILTERED, ppt_list),) + (() if only_filtered else ((KERNEL_EXACT, (0,)),)):
            for ppt in ppts:
                p = lib.default_params(kernel=kernel, pixels_per_thread=ppt)
                ms = []
                for it in range(4):
                    ctx.render_lbuffer(p, extras=False, out=dict(L=L))
                    ms.append(ctx.timing().render_ms)
                best = min(ms[1:])
                ctx.count_exact(True)
                ctx.render_lbuffer(p, extras=False, out=dict(L=L))
                t = ctx.timing()
                ctx.count_exact(False)
                tests = size * size * nt
                print(json.dumps({"size": size, "kernel": kernel, "ppt": ppt, "ms": best, "all_ms": ms,
                                  "tests_per_s": tests / (best * 1e-3), "alg_tflops": tests * 19 / (best * 1e-3) / 1e12,
                                  "exact_frac": t.exact_tests / tests}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
