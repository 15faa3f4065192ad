#!/usr/bin/env python3
"""Golden vectors of the shaded variant (src/main-omp.cxx semantics) from the REFERENCE's own classes
(oracle/_ref, build container only): dragon + textured background quad at 128 x 128, once with the
reference's Bangor_Logo_A1.jpg texture (all white in the shipped file) and once with a seeded noise
texture that exercises the texel lookup. z / tri_id / shadow come from the oracle (bit-identical rgb)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import Oracle, Reference  # noqa: E402
from simpleraytracing_b200 import scene  # noqa: E402

ref, orc = Reference(), Oracle()
out = {}
for name, tex in (("logo", None), ("noise", scene.noise_texture())):
    meshes = scene.shaded_scene(texture=tex)
    cam = ref.camera(meshes[0]["tri"], [len(meshes[0]["tri"])])
    rgb = ref.render_shaded(meshes, cam, 128, 128)
    o = orc.render_shaded(meshes, cam, 128, 128)
    assert (o["rgb"] == rgb).all()
    out[f"rgb_{name}"] = rgb
    out["z"], out["tri_id"], out["shadow"] = o["z"], o["tri_id"], o["shadow"]
    print(name, "hit px", int((o["tri_id"] >= 0).sum()), "dragon px", int(((o["tri_id"] >= 0) & (o["tri_id"] < 22866)).sum()),
          "shadowed", int(o["shadow"].sum()), "colours", len(np.unique(rgb.reshape(-1, 3), axis=0)))
np.savez_compressed(ROOT / "tests" / "golden" / "shaded_dragon_logo_128.npz", **out)
