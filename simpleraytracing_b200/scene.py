"""Scene helpers for the tests and bench.py (harness code, not the hot path).

Builds the inputs BASELINE.json's configs name: the dragon mesh (fixture
tests/golden/dragon_mesh.npz, generated from the reference's dragon.ply by
tools/make_fixtures.py), the textured logo quad (createBackground, src/main-omp.cxx:525-555)
and the synthetic UV-sphere of config C5.
"""
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"

MU_MESH0 = np.float32(0.1037)     # "soft tissue", main-pthreads-lbuffer.cxx:800
MU_OTHER = np.float32(0.3971)     # "bone", :802


def soup(vertices, faces):
    """Indexed mesh -> triangle soup [n, 9] (TriangleMesh::setGeometry(vertices, indices),
    src/TriangleMesh.cxx:104-131: a pure gather, no arithmetic)."""
    v = np.asarray(vertices, np.float32).reshape(-1, 3)
    f = np.asarray(faces).reshape(-1, 3)
    return np.ascontiguousarray(v[f.reshape(-1)].reshape(-1, 9))


def dragon():
    d = np.load(GOLDEN / "dragon_mesh.npz")
    return soup(d["vertices"], d["faces"])


def logo_quad(lo, hi):
    """createBackground (src/main-omp.cxx:525-555 == main-pthreads-lbuffer.cxx:670-692):
    float arithmetic `upper + range * 0.1f` etc. on the scene bbox."""
    lo = np.asarray(lo, np.float32); hi = np.asarray(hi, np.float32)
    rng = hi - lo
    f = np.float32
    x = hi[0] + rng[0] * f(0.1)
    y0 = lo[1] - rng[1] * f(0.5); y1 = hi[1] + rng[1] * f(0.5)
    z0 = lo[2] - rng[2] * f(0.5); z1 = hi[2] + rng[2] * f(0.5)
    verts = np.array([[x, y0, z0], [x, y1, z0], [x, y1, z1], [x, y0, z1]], np.float32)
    faces = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    return soup(verts, faces)


def uv_sphere(n_lat=500, n_lon=1000, radius=100.0, centre=(0.0, -670.0, 300.0), noise=0.05, seed=1234):
    """Closed UV-sphere with radial noise, outward winding (config C5: 998 000 triangles)."""
    rng = np.random.default_rng(seed)
    th = np.linspace(0.0, np.pi, n_lat + 1)[1:-1]                    # interior latitudes
    ph = np.linspace(0.0, 2.0 * np.pi, n_lon, endpoint=False)
    r = radius * (1.0 + noise * rng.standard_normal((len(th), n_lon)).clip(-2, 2) * 0.5)
    x = r * np.sin(th)[:, None] * np.cos(ph)[None, :]
    y = r * np.sin(th)[:, None] * np.sin(ph)[None, :]
    z = r * np.cos(th)[:, None] * np.ones_like(ph)[None, :]
    ring = np.stack([x, y, z], -1).reshape(-1, 3)
    verts = np.concatenate([[[0, 0, radius]], ring, [[0, 0, -radius]]]) + np.asarray(centre)
    nr = len(th)
    idx = lambda i, j: 1 + i * n_lon + (j % n_lon)
    j = np.arange(n_lon)
    faces = [np.stack([np.zeros(n_lon, int), idx(0, j), idx(0, j + 1)], 1)]
    for i in range(nr - 1):
        a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
        faces.append(np.stack([a, b, c], 1)); faces.append(np.stack([a, c, d], 1))
    south = len(verts) - 1
    faces.append(np.stack([np.full(n_lon, south), idx(nr - 1, j + 1), idx(nr - 1, j)], 1))
    return soup(verts.astype(np.float32), np.concatenate(faces).astype(np.uint32))
