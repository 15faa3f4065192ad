#!/usr/bin/env python3
"""One full frame + share 0 of 8 of the filtered kernel for a box shape (target of ncu captures). usage: one_share.py [PXxPYxNB] [size]"""
import os, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simpleraytracing_b200 import Context, Library, scene, KERNEL_FILTERED  # noqa: E402
from simpleraytracing_b200.capi import box_shape  # noqa: E402
sh = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "16x2x4").split("x"))
size = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
lib = Library(os.environ.get("SRT_LIB")); ctx = Context(0, lib)
tri = scene.dragon(); lo, hi = lib.scene_bbox([(tri, 0.1037)])
ctx.set_meshes([(tri, 0.1037)]); ctx.set_camera(lib.camera_from_bbox(lo, hi, size, size))
frame = torch.zeros(size * size, dtype=torch.float32, device="cuda")
p = lib.default_params(kernel=KERNEL_FILTERED, pixels_per_thread=box_shape(*sh))
for _ in range(2):
    ctx.render_lbuffer_device(p, (0, size), frame.data_ptr()); print("full", ctx.timing().render_ms)
    ctx.render_lbuffer_interleaved_device(p, 8, 0, frame.data_ptr()); print("share 0/8", ctx.timing().render_ms)
    ctx.render_lbuffer_interleaved_device(p, 8, 7, frame.data_ptr()); print("share 7/8", ctx.timing().render_ms)
ctx.close()
