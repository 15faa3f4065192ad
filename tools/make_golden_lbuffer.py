#!/usr/bin/env python3
"""Generate L-buffer golden vectors with the REFERENCE's own classes (oracle/_ref).

Runs only in the build container (needs /root/reference to build oracle/_ref/libsrt_ref.so).
For each size the dragon is rendered through ref_render_lbuffer -- i.e. the reference's
Vec3/Ray/Triangle/TriangleMesh code driven by the restated main-pthreads-lbuffer.cxx loop --
and the per-pixel results are stored compactly:

  tests/golden/lbuffer_dragon_<N>.npz   L (float32, pre-repair, -1 = flagged), dist (float32),
                                        ssum (int8), hits (uint8); all N*N, row-major
  tests/golden/lbuffer_dragon_logo_128.npz   same for the 2-mesh scene dragon + logo quad
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import Reference  # noqa: E402
from simpleraytracing_b200 import scene  # noqa: E402

OUT = ROOT / "tests" / "golden"


def save(name, r):
    assert np.abs(r["ssum"]).max() < 127 and r["hits"].max() < 255
    np.savez_compressed(OUT / name, L=r["L"], dist=r["dist"], ssum=r["ssum"].astype(np.int8),
                        hits=r["hits"].astype(np.uint8))
    print(name, "flagged", int((r["L"] == -1).sum()), "hit pixels", int((r["hits"] > 0).sum()),
          "hits", int(r["hits"].sum()))


def main():
    ref = Reference()
    tri = scene.dragon()
    cam = ref.camera(tri, [len(tri)])
    for n in (128, 256, 512):
        save(f"lbuffer_dragon_{n}.npz", ref.render_lbuffer(tri, [len(tri)], cam, n, n))
    # two meshes: the camera comes from the dragon bbox (the quad is derived from it, then appended)
    quad = scene.logo_quad(cam.vec("lower"), cam.vec("upper"))
    both = np.concatenate([tri, quad])
    save("lbuffer_dragon_logo_128.npz", ref.render_lbuffer(both, [len(tri), len(quad)], cam, 128, 128))


if __name__ == "__main__":
    main()
