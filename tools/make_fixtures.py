#!/usr/bin/env python3
"""Generate the committed test fixtures under tests/golden/ from the reference's assets.

Runs ONLY in the build container (it reads /root/reference, which does not exist
on the GPU box). Outputs are small binary fixtures:

  tests/golden/dragon_mesh.npz          vertices (float32 [11429,3]) + faces (uint32 [22866,3])
                                        of /root/reference/dragon.ply, file order preserved
                                        (the reference loads it through Assimp, which keeps
                                        face order for this file -- SURVEY.md section 4).
  tests/golden/dragon_128_serial.npy    the 128x128 float values printed in
                                        /root/reference/out/dragon-128x128-serial.txt
                                        (output of src/main.cxx, 6 significant digits),
                                        parsed to float64 so that '%.6g' reproduces the
                                        original strings exactly.
  tests/golden/logo_texture.npz         Bangor_Logo_A1.jpg decoded once to raw RGB (uint8)
                                        -- texture of the shaded variant's background quad
                                        (src/main-omp.cxx:551).
"""
import struct
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def read_binary_ply(path: Path):
    raw = path.read_bytes()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    header = raw[:end].decode("ascii").splitlines()
    assert "format binary_little_endian 1.0" in header
    nv = nf = None
    for line in header:
        tok = line.split()
        if tok[:2] == ["element", "vertex"]:
            nv = int(tok[2])
        if tok[:2] == ["element", "face"]:
            nf = int(tok[2])
    verts = np.frombuffer(raw, dtype="<f4", count=nv * 3, offset=end).reshape(nv, 3).copy()
    off = end + nv * 12
    faces = np.empty((nf, 3), dtype=np.uint32)
    for i in range(nf):
        n = raw[off]
        assert n == 3, "non-triangle face"
        faces[i] = struct.unpack_from("<3i", raw, off + 1)
        off += 13
    assert off == len(raw)
    return verts, faces


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    verts, faces = read_binary_ply(REF / "dragon.ply")
    np.savez_compressed(OUT / "dragon_mesh.npz", vertices=verts, faces=faces)
    print("dragon:", verts.shape, faces.shape, "lo", verts.min(0), "hi", verts.max(0))

    rows = (REF / "out" / "dragon-128x128-serial.txt").read_text().split("\n")
    vals = np.array([[float(t) for t in r.split("\t")] for r in rows], dtype=np.float64)
    assert vals.shape == (128, 128)
    # the fixture must regenerate the original strings
    for r, row in enumerate(rows):
        for c, tok in enumerate(row.split("\t")):
            assert "%.6g" % vals[r, c] == tok, (r, c, tok)
    np.save(OUT / "dragon_128_serial.npy", vals)
    print("golden:", vals.shape, "min", vals.min(), "n<80", int((vals < 80).sum()))

    try:
        from PIL import Image
        im = np.asarray(Image.open(REF / "Bangor_Logo_A1.jpg").convert("RGB"))
        np.savez_compressed(OUT / "logo_texture.npz", rgb=im)
        print("texture:", im.shape)
    except Exception as e:  # texture only matters for the shaded variant
        print("texture skipped:", e, file=sys.stderr)


if __name__ == "__main__":
    main()
