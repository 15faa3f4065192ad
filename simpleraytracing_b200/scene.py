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


# ---- shaded variant (src/main-omp.cxx) ---------------------------------------------------------------
DRAGON_MATERIAL = ((np.float32(np.float64(np.float32(1)) * 0.2), 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0), 1.0)   # :197
DEFAULT_MATERIAL = ((1.0, 1.0, 1.0), (1.0, 1.0, 1.0), (1.0, 1.0, 1.0), 1.0)                # Material(), Material.inl:76-83

# texture coordinates of the background quad, per vertex (src/main-omp.cxx:537-542), expanded per triangle
_QUAD_TC = np.array([[0, 1, 0], [1, 1, 0], [1, 0, 0], [0, 0, 0]], np.float32)
QUAD_TEXCOORDS = np.ascontiguousarray(_QUAD_TC[np.array([0, 1, 2, 0, 2, 3])].reshape(2, 9))


def logo_texture():
    """Bangor_Logo_A1.jpg decoded to raw RGB by tools/make_fixtures.py (1240 x 1754 x 3)."""
    return np.load(GOLDEN / "logo_texture.npz")["rgb"]


def shaded_scene(texture=None):
    """Config C1: dragon (material of main-omp.cxx:197) + textured background quad derived from the dragon bbox."""
    tri = dragon()
    lo = tri.reshape(-1, 3).min(0); hi = tri.reshape(-1, 3).max(0)
    quad = logo_quad(lo, hi)
    return [dict(tri=tri, texcoords=None, material=DRAGON_MATERIAL, texture=None),
            dict(tri=quad, texcoords=QUAD_TEXCOORDS, material=DEFAULT_MATERIAL,
                 texture=logo_texture() if texture is None else texture)]


def shaded_camera(lib, meshes, W, H):
    """Camera of the shaded drivers: origin/detector/up/right from the bbox of mesh 0 only (main-omp.cxx:208-241
    runs before the background quad is appended), pixel spacing from the bbox of ALL meshes (renderLoop, :594-607),
    light at origin + up * 100.0 (:235).  Returns (camera, light_position)."""
    tri0 = np.asarray(meshes[0]["tri"], np.float32)
    lo, hi = lib.scene_bbox([(tri0, 0.0)])
    cam = lib.camera_from_bbox(lo, hi, W, H)
    lo_all, hi_all = lib.scene_bbox([(np.asarray(m["tri"], np.float32), 0.0) for m in meshes])
    rng = hi_all - lo_all                                            # float32
    res1 = rng[2] / np.float32(W); res2 = rng[1] / np.float32(H)
    sp = np.float32(2) * max(res1, res2)
    cam.pixel_spacing[0] = cam.pixel_spacing[1] = float(sp)
    up = cam.vec("up"); origin = cam.vec("origin")
    light = origin + (up.astype(np.float64) * 100.0).astype(np.float32)   # Vec3::operator*(double), then float add
    return cam, light


def noise_texture(h=64, w=96, seed=1):
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, 3), dtype=np.uint8)
