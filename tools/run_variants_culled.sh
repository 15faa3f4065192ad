#!/bin/bash
# usage: tools/run_variants_culled.sh name1 name2 ...  -- 2048^2 dragon frame + rank-0-of-8 share with the culled kernel per library build
for v in "$@"; do echo "== $v"; SRT_LIB=simpleraytracing_b200/lib/variants/libsrt_b200_$v.so python - <<PY
import os, sys
sys.path.insert(0, ".")
import torch
from simpleraytracing_b200 import Context, Library, scene
lib = Library(os.environ["SRT_LIB"]); ctx = Context(0, lib)
tri = scene.dragon(); lo, hi = lib.scene_bbox([(tri, 0.1037)])
ctx.set_meshes([(tri, 0.1037)]); ctx.set_camera(lib.camera_from_bbox(lo, hi, 2048, 2048))
buf = torch.zeros(2048 * 2048, dtype=torch.float32, device="cuda")
sr = lib.interleave_slot_rows(2048, 8); slot = torch.zeros(sr * 2048, dtype=torch.float32, device="cuda")
p = lib.default_params(kernel=4)
for name, fn in (("full", lambda: ctx.render_lbuffer_device(p, (0, 2048), buf.data_ptr())),
                 ("part0of8", lambda: ctx.render_lbuffer_interleaved_device(p, 8, 0, slot.data_ptr()))):
    ms = []
    for _ in range(6):
        fn(); ms.append(ctx.timing().render_ms)
    print(name, "ms", round(min(ms[1:]), 4), flush=True)
ctx.close()
PY
done
