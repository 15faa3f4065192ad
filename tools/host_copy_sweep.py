"""Frames returned into a pageable host buffer (srt_render_xray, csrc/host_copy.h): wall time of the plugin call
srt_set_meshes + srt_set_camera + srt_render_xray on the 2048^2 dragon frame for every combination of helper threads,
DMA chunk size and polling time of the copy threads, next to the driver's own pageable copy (threads = 0) and a
page-locked buffer.  python tools/host_copy_sweep.py > profiles/r02_host_copy_sweep.txt"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from simpleraytracing_b200 import Context, Library, scene          # noqa: E402
from simpleraytracing_b200.capi import PinnedArray                  # noqa: E402


def main():
    lib = Library()
    ctx = Context(0, lib)
    tri = scene.dragon()
    W = H = 2048
    meshes = [(tri, float(scene.MU_MESH0))]
    lo, hi = lib.scene_bbox(meshes)
    cam = lib.camera_from_bbox(lo, hi, W, H)
    p = lib.default_params(kernel=2)
    img = np.zeros(W * H, np.float32)
    pin = PinnedArray(lib, (W * H,), np.float32)

    def call(out):
        ctx.set_meshes(meshes); ctx.set_camera(cam); ctx.render_xray(p, out=out)

    def timed(out, n=12):
        for _ in range(3):
            call(out)
        ts = []
        for _ in range(n):
            t0 = time.perf_counter(); call(out); ts.append(time.perf_counter() - t0)
        return float(np.median(ts)) * 1e3, float(np.min(ts)) * 1e3

    print(f"plugin call on the {W}x{H} dragon frame, median (best) ms of 12 calls")
    print(f"page-locked buffer                                   : {'%.3f (%.3f)' % timed(pin.array)}")
    ctx.set_host_copy_threads(0)
    ref = img.copy(); call(ref)
    print(f"pageable, driver's copy (threads 0)                  : {'%.3f (%.3f)' % timed(img)}")
    for spin in (0, 2000):
        for chunk_kb in (256, 512, 1024, 2048, 4096):
            row = []
            for threads in (2, 3, 4, 5, 6):
                ctx.set_host_copy_threads(threads); ctx.tune_host_copy(chunk_kb << 10, spin)
                m, b = timed(img)
                assert np.array_equal(img.view(np.uint32), ref.view(np.uint32))
                row.append(f"{threads}: {m:.3f} ({b:.3f})")
            print(f"pageable, staged, poll {spin:4d} us, chunk {chunk_kb:4d} KiB, threads -> " + "  ".join(row))
    ctx.set_host_copy_threads(-1); ctx.tune_host_copy(0, -1)
    print(f"pageable, staged, library defaults                   : {'%.3f (%.3f)' % timed(img)}")
    ctx.close()


if __name__ == "__main__":
    main()
