#!/bin/bash
for v in "$@"; do echo "== $v"; SRT_LIB=simpleraytracing_b200/lib/variants/libsrt_b200_$v.so python - <<PY
import os, sys
sys.path.insert(0, ".")
import torch
from simpleraytracing_b200 import Context, Library, scene
lib = Library(os.environ["SRT_LIB"]); ctx = Context(0, lib)
tri = scene.dragon(); lo, hi = lib.scene_bbox([(tri, 0.1037)])
ctx.set_meshes([(tri, 0.1037)]); ctx.set_camera(lib.camera_from_bbox(lo, hi, 2048, 2048))
for n_parts in (8, 4):
    sr = lib.interleave_slot_rows(2048, n_parts)
    slot = torch.zeros(sr * 2048, dtype=torch.float32, device="cuda")
    for ppt in (8, 7):
        p = lib.default_params(kernel=2, pixels_per_thread=ppt)
        ms = []
        for _ in range(4):
            ctx.render_lbuffer_interleaved_device(p, n_parts, 0, slot.data_ptr()); ms.append(ctx.timing().render_ms)
        print("parts", n_parts, "ppt", ppt, "ms", round(min(ms[1:]), 4), flush=True)
ctx.close()
PY
done
