import sys, time
sys.path.insert(0, "/root/repo")
import torch
from simpleraytracing_b200 import Context, Library, scene
lib = Library(); ctx = Context(0, lib)
tri = scene.uv_sphere()
lo, hi = lib.scene_bbox([(tri, 0.1)])
ctx.set_meshes([(tri, 0.1037)])
for size in (512, 1024, 2048):
    ctx.set_camera(lib.camera_from_bbox(lo, hi, size, size))
    L = torch.zeros(size * size, dtype=torch.float32, device="cuda")
    p = lib.default_params(kernel=2)
    for rep in range(2):
        t0 = time.time()
        ctx.render_lbuffer_device(p, (0, size), L.data_ptr()); t = ctx.timing()
        print(size, rep, f"pass {t.render_ms:.2f} ms scan {t.main_kernel_ms:.2f} ms cols {t.scan_cols} wall {time.time()-t0:.2f}s", flush=True)
