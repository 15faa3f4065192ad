#!/usr/bin/env python3
"""Development sweeps on ONE GPU (device time from the library's CUDA events, best of N after a warm-up).

  python tools/sweep.py boxes  [size]          every box shape of the filtered kernel: full frame + the 8 block-cyclic shares
  python tools/sweep.py splits [size] [shape]  full frame and share 0 of 8 with forced box-column counts per CTA
  python tools/sweep.py parts  [size] [shape]  per-share kernel time for N = 2, 4, 8 (load balance across ranks)
  python tools/sweep.py kernels [size] [scene] one frame with every kernel selector (+ exact-test counts)

shape = PXxPYxNB, e.g. 16x2x4 (box width x box height x triangles per thread).  Output: one line per measurement; copy what matters into profiles/.
"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simpleraytracing_b200 import Context, Library, scene, KERNEL_FILTERED, KERNEL_CULLED, KERNEL_EXACT  # noqa: E402
from simpleraytracing_b200.capi import box_shape  # noqa: E402

SHAPES = [(16, 2, 4), (32, 2, 4), (16, 4, 4), (8, 8, 4), (32, 4, 4), (32, 8, 4), (16, 16, 4), (32, 16, 4), (64, 8, 4), (64, 16, 4)]


def parse_shape(s):
    return tuple(int(v) for v in s.split("x"))


def best(fn, ctx, n=4):
    ms = []
    for _ in range(n):
        fn()
        ms.append(ctx.timing().render_ms)
    return min(ms[1:])


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "boxes"
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    lib = Library(os.environ.get("SRT_LIB")); ctx = Context(0, lib)      # SRT_LIB: a tuning variant of the library
    which = sys.argv[3] if mode == "kernels" and len(sys.argv) > 3 else "dragon"
    tri = scene.dragon() if which == "dragon" else scene.uv_sphere()
    lo, hi = lib.scene_bbox([(tri, 0.1037)])
    ctx.set_meshes([(tri, 0.1037)]); ctx.set_camera(lib.camera_from_bbox(lo, hi, size, size))
    frame = torch.zeros(size * size, dtype=torch.float32, device="cuda")

    def full(p):
        return best(lambda: ctx.render_lbuffer_device(p, (0, size), frame.data_ptr()), ctx)

    def share(p, n_parts, part):
        return best(lambda: ctx.render_lbuffer_interleaved_device(p, n_parts, part, frame.data_ptr()), ctx)

    if mode == "boxes":
        for sh in SHAPES:
            p = lib.default_params(kernel=KERNEL_FILTERED, pixels_per_thread=box_shape(*sh))
            f = full(p)
            parts = [round(share(p, 8, r), 4) for r in range(8)]
            print(f"box {sh} {size}^2: full {f:.4f} ms | shares of 8: {parts} sum {sum(parts):.3f} max {max(parts):.4f} "
                  f"-> efficiency bound {f / 8 / max(parts):.3f}", flush=True)
    elif mode == "splits":
        sh = parse_shape(sys.argv[3]) if len(sys.argv) > 3 else (16, 2, 4)
        p = lib.default_params(kernel=KERNEL_FILTERED, pixels_per_thread=box_shape(*sh))
        for zs in (0, 4, 8, 16, 32, 64, 128):
            ctx.force_scan_cols(zs)
            print(f"box {sh} box columns per CTA {zs}: full {full(p):.4f} ms, share 0/8 {share(p, 8, 0):.4f} ms, share 0/4 {share(p, 4, 0):.4f} ms", flush=True)
        ctx.force_scan_cols(0)
    elif mode == "parts":
        sh = parse_shape(sys.argv[3]) if len(sys.argv) > 3 else (16, 2, 4)
        for kern in (KERNEL_FILTERED, KERNEL_CULLED):
            p = lib.default_params(kernel=kern, pixels_per_thread=box_shape(*sh))
            f = full(p)
            for n in (2, 4, 8):
                parts = [round(share(p, n, r), 4) for r in range(n)]
                print(f"kernel {kern} box {sh} N={n}: full {f:.4f} | {parts} max {max(parts):.4f} -> bound {f / n / max(parts):.3f}", flush=True)
    elif mode == "kernels":
        ctx.count_exact(True)
        for k in (KERNEL_CULLED, KERNEL_FILTERED) + ((KERNEL_EXACT,) if size <= 1024 else ()):
            p = lib.default_params(kernel=k)
            ms = full(p)
            t = ctx.timing()
            print(f"kernel {k} {which} {size}^2: {ms:.4f} ms, exact tests {t.exact_tests}, pairs {t.pair_tests}, "
                  f"{t.pair_tests / ms / 1e9:.2f} Gtests/ms", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
