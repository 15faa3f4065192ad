"""Host-side logic of the product (camera, bbox, partition, pre-pack, filter margins) against the oracle.
CPU only: these entry points of libsrt_b200.so do no GPU work."""
import numpy as np
import pytest

from conftest import bits
from simpleraytracing_b200 import scene


def test_bbox_and_camera_bit_exact(srtlib, oracle, dragon):
    lo, hi = srtlib.scene_bbox([(dragon, 0.1037)])
    cam_o = oracle.camera(dragon)
    assert (bits(lo) == bits(cam_o.vec("lower"))).all() and (bits(hi) == bits(cam_o.vec("upper"))).all()
    for W, H in ((128, 128), (1024, 1024), (2048, 2048), (1512, 1512), (640, 480), (3, 7)):
        cam = srtlib.camera_from_bbox(lo, hi, W, H)
        for n in ("origin", "detector", "up", "right"):
            assert (bits(cam.vec(n)) == bits(cam_o.vec(n))).all(), (n, W, H)
        sp = oracle.pixel_spacing(cam_o, W, H)
        assert bits(np.float32(cam.pixel_spacing[0])) == bits(sp) and cam.pixel_spacing[1] == cam.pixel_spacing[0]


def test_bbox_multi_mesh(srtlib, oracle, dragon):
    lo, hi = srtlib.scene_bbox([(dragon, 0.1)])
    quad = scene.logo_quad(lo, hi)
    lo2, hi2 = srtlib.scene_bbox([(dragon, 0.1), (quad, 0.3)])
    cam = oracle.camera(np.concatenate([dragon, quad]))
    assert (bits(lo2) == bits(cam.vec("lower"))).all() and (bits(hi2) == bits(cam.vec("upper"))).all()


def test_camera_rejects_bad_arguments(srtlib):
    from simpleraytracing_b200 import SrtError
    with pytest.raises(SrtError):
        srtlib.camera_from_bbox([0, 0, 0], [1, 1, 1], 0, 8)


def test_partition_rows_reference_rule(srtlib, oracle):
    for H, n in ((2048, 8), (2048, 3), (10, 4), (7, 8), (1, 1), (1512, 5)):
        s, e = oracle.partition(H, n)
        got = [srtlib.partition_rows(H, n, i) for i in range(n)]
        for i in range(n):
            if e[i] + 1 > s[i] or H >= n:       # the reference's [start, end] is inclusive
                assert got[i] == (int(s[i]), int(e[i]) + 1), (H, n, i)
        assert got[0][0] == 0 and got[-1][1] == H
        assert all(got[i][1] == got[i + 1][0] for i in range(n - 1))


def test_triangle_normals_bit_exact(srtlib, oracle, dragon):
    n = srtlib.triangle_normals(dragon[:4000])
    ref = np.array([oracle.triangle_normal(t) for t in dragon[:4000]])
    assert (bits(n) == bits(ref)).all()


def test_prepack_exact_constants(srtlib, oracle, dragon):
    """ExactTri must hold the floats Ray::intersect would compute (Ray.cxx:86-122), bit for bit."""
    lo, hi = srtlib.scene_bbox([(dragon, 0.1)])
    cam = srtlib.camera_from_bbox(lo, hi, 256, 256)
    pk = srtlib.debug_prepack(dragon, cam)
    f = np.float32
    p1, p2, p3 = dragon[:, 0:3], dragon[:, 3:6], dragon[:, 6:9]
    o = cam.vec("origin")
    e1 = p2 - p1; e2 = p3 - p1; tv = o[None, :] - p1

    def cross(a, b):      # Vec3.inl:321-329, float32 ops, one rounding each
        return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                         a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], 1)
    qv = cross(tv, e1)
    e2q = (e2[:, 0] * qv[:, 0] + e2[:, 1] * qv[:, 1]) + e2[:, 2] * qv[:, 2]
    E = pk["exact"]
    expect = np.concatenate([e2, e1, tv, qv, e2q[:, None], srtlib.triangle_normals(dragon)], 1).astype(f)
    assert expect.dtype == f and (bits(E) == bits(expect)).all()
    pad = srtlib.pack_info()[0]
    assert pk["unfilterable"] == 0 and pk["n_padded"] == (len(dragon) + pad - 1) // pad * pad


BOXES = [(16, 2), (16, 1), (32, 1), (8, 4), (8, 1), (32, 2), (16, 4), (8, 8)]          # box shapes the filtered kernel is built for


@pytest.mark.parametrize("size,scale", [(128, 1.0), (2048, 1.0), (640, 3000.0), (640, 1e-4)])
def test_level0_box_bound(srtlib, dragon, size, scale):
    """Word 7 of the level-0 record is the box bound R of the filtered kernel's gate |Y(c)| - |D(c)| > R: it must
    cover the box radius of Y and D plus twice the strip margin (filter proof item 7), whatever the scene scale."""
    tri = (dragon * np.float32(scale)).astype(np.float32)
    lo, hi = srtlib.scene_bbox([(tri, 0.1)])
    cam = srtlib.camera_from_bbox(lo, hi, size, size)
    for bw, bh in BOXES:
        pk = srtlib.debug_prepack(tri, cam, box=(bw, bh))
        assert pk["unfilterable"] == 0
        f = pk["filt"][:len(tri)].astype(np.float64)
        hx, hy = (bw - 1) / 2.0, (bh - 1) / 2.0
        need = hx * (np.abs(f[:, 0]) + np.abs(f[:, 3])) + hy * (np.abs(f[:, 1]) + np.abs(f[:, 4])) + 2.0 * f[:, 6]
        assert (f[:, 7] >= need).all() and (f[:, 7] <= need * (1 + 1e-5) + 1e-30).all()
        g = pk["filt2"][:len(tri)].astype(np.float64)          # second strip: same coefficients as the level-1 record
        assert (pk["filt"][:len(tri), 8:11] == pk["filt2"][:len(tri), 0:3]).all()
        need2 = hx * (np.abs(f[:, 8]) + np.abs(f[:, 3])) + hy * (np.abs(f[:, 9]) + np.abs(f[:, 4])) + 2.0 * g[:, 3]
        assert (f[:, 11] >= need2).all() and (f[:, 11] <= need2 * (1 + 1e-5) + 1e-30).all()
        assert np.isfinite(pk["filt"]).all() and np.isfinite(pk["filt2"]).all() and (f[:, 6] > 0).all()
    # padding records never survive: Y = 3e38, D = 0, R = 0
    pad = pk["filt"][len(tri):]
    assert len(pad) == pk["n_padded"] - len(tri) and (pad[:, 2] > 1e38).all() and (pad[:, 7] == 0).all() and (pad[:, 11] == 0).all()


@pytest.mark.parametrize("size", [48, 96])
def test_filter_never_rejects_a_hit_dragon(srtlib, oracle, dragon, size):
    """CPU emulation of the kernels' FFMA filter vs the oracle for every (pixel, triangle) pair, every box shape."""
    lo, hi = srtlib.scene_bbox([(dragon, 0.1)])
    cam = srtlib.camera_from_bbox(lo, hi, size, size)
    for box in BOXES + [(32, 8), (64, 16)] if size == 48 else [BOXES[0], (64, 16)]:      # (64, 16): the default scan box
        pk = srtlib.debug_prepack(dragon, cam, box=box)
        a = oracle.filter_audit(dragon, oracle.camera(dragon), size, size, pk)
        assert a["pairs"] == size * size * len(dragon)
        assert a["violations"] == 0 and a["box3_violations"] == 0 and a["box_violations"] == 0, (box, a)
        assert a["hits"] > 0 and a["survive1"] >= a["hits"]
        if box in BOXES:       # the gate must actually filter (a box is up to 2/3 of a 48-pixel frame wide; the big boxes exceed it)
            assert a["survive0_box"] < (0.25 if size == 48 else 0.02) * a["pairs"]
        assert a["survive1"] < 10 * a["hits"]


def _adversarial_mesh(cam_o, oracle, W, H, rng):
    """Triangles whose vertices / edges lie exactly on pixel rays (the boundary cases of u, v, u+v)."""
    o = cam_o.vec("origin").astype(np.float64)
    tris = []
    for _ in range(300):
        pix = rng.integers(0, [H, W], size=(3, 2))
        ts = rng.uniform(150, 300, size=3)
        pts = []
        for (r, c), t in zip(pix, ts):
            _, db = oracle.make_ray(cam_o, W, H, int(r), int(c))
            pts.append(o + t * db.astype(np.float64))
        kind = rng.integers(0, 4)
        if kind == 1:      # sliver: third vertex on the edge of the first two
            pts[2] = pts[0] + (pts[1] - pts[0]) * rng.uniform()
        elif kind == 2:    # tiny triangle around one pixel ray
            pts[1] = pts[0] + rng.normal(size=3) * 1e-3
            pts[2] = pts[0] + rng.normal(size=3) * 1e-3
        elif kind == 3:    # huge triangle
            pts[1] = pts[0] + rng.normal(size=3) * 500
            pts[2] = pts[0] + rng.normal(size=3) * 500
        tris.append(np.concatenate(pts))
    return np.asarray(tris, np.float32)


def test_box_tests_never_reject_a_hit_ragged_frame(srtlib, oracle, dragon):
    """The box tests of the filtered kernel (every box shape) and of the culled kernel (bin 256 x 128, CTA tile 64 x 16,
    warp tile 32 x 8, run 8 x 1; boxes clamped to the frame) emulated on the CPU against the oracle, on a frame whose
    size is not a multiple of any box size and which is larger than one bin in both directions."""
    W, H = 300, 150
    lo, hi = srtlib.scene_bbox([(dragon, 0.1)])
    cam = srtlib.camera_from_bbox(lo, hi, W, H)
    tri = np.ascontiguousarray(dragon[::8])
    for box in BOXES:
        pk = srtlib.debug_prepack(tri, cam, box=box)
        a = oracle.filter_audit(tri, oracle.camera(dragon), W, H, pk)
        assert a["hits"] > 1500 and a["box_violations"] == 0 and a["violations"] == 0 and a["box3_violations"] == 0, (box, a)


def test_filter_never_rejects_a_hit_adversarial(srtlib, oracle, dragon):
    rng = np.random.default_rng(11)
    cam_o = oracle.camera(dragon)
    lo, hi = cam_o.vec("lower"), cam_o.vec("upper")
    W = H = 40
    tri = _adversarial_mesh(cam_o, oracle, W, H, rng)
    cam = srtlib.camera_from_bbox(lo, hi, W, H)
    for box in BOXES:
        pk = srtlib.debug_prepack(tri, cam, box=box)
        a = oracle.filter_audit(tri, cam_o, W, H, pk)
        assert a["violations"] == 0 and a["box3_violations"] == 0 and a["box_violations"] == 0 and a["hits"] > 500, (box, a)


ALL_BOXES = BOXES + [(32, 4), (32, 8), (32, 16), (64, 8), (64, 16), (16, 16)]           # + the big-box shapes (default: 64 x 16)


def _random_soup(cam_vecs, rng, n, scale):
    """Triangles scattered in front of a camera: sizes over seven decades, slivers, needles (the generator of the GPU
    property test, tests/test_gpu_parity.py)."""
    o, det, up, right, sp, W, H = cam_vecs
    tris = np.empty((n, 9), np.float64)
    for i in range(n):
        u = (rng.uniform(0, W) - W / 2.0) * sp; v = (rng.uniform(0, H) - H / 2.0) * sp
        d = det + up * v + right * u - o
        d /= np.linalg.norm(d)
        p = o + d * rng.uniform(0.3, 1.2) * np.linalg.norm(det - o)
        size = scale * 10.0 ** rng.uniform(-3, 4) * 1e-2
        kind = rng.integers(0, 4)
        a = rng.normal(size=3) * size; b = rng.normal(size=3) * size
        if kind == 1:
            b = a * rng.uniform(0.2, 0.9) + rng.normal(size=3) * size * 1e-4        # sliver
        elif kind == 2:
            b = rng.normal(size=3) * size * 1e-3                                    # needle
        tris[i] = np.concatenate([p, p + a, p + b])
    return np.ascontiguousarray(tris, np.float32)


@pytest.mark.parametrize("frame,scale,seed", [((96, 72), 1.0, 1), ((200, 120), 4096.0, 2), ((160, 160), 1e-3, 3),
                                              ((2048, 1536), 1.0, 4), ((8192, 8192), 2500.0, 5), ((8192, 8192), 1.0, 6)])
def test_filter_property_random_soups_and_cameras_cpu(srtlib, oracle, frame, scale, seed):
    """The filter proof's evidence without a GPU: random triangle soups -- sizes over seven decades,
    slivers, needles, scene scales 1e-3 ... 4096 -- under random general (not exactly orthonormal) cameras; every gate of
    the kernels emulated with fmaf next to the oracle's intersect for every pair, every box shape including the default
    64 x 16.  The large frames (large |x|, |y| enter every margin) are audited on rows at the top edge, in the middle and at
    the bottom edge."""
    from simpleraytracing_b200 import Camera
    W, H = frame
    rng = np.random.default_rng(seed)
    centre = rng.normal(size=3) * 100.0 * scale
    d = rng.normal(size=3); d /= np.linalg.norm(d)
    dist_o, dist_d = rng.uniform(180, 400) * scale, rng.uniform(60, 200) * scale
    origin = centre - d * dist_o; detector = centre + d * dist_d
    a = np.cross(d, rng.normal(size=3)); a /= np.linalg.norm(a)
    b = np.cross(d, a)
    up = a + 0.1 * rng.normal(size=3); right = b + 0.1 * rng.normal(size=3)
    sp = rng.uniform(0.15, 0.3) * scale * 1024.0 / max(W, H)
    # the oracle derives the pixel spacing from the bbox range (lbuffer.cxx:739-741): choose a range that yields ~sp and
    # hand the product the float the oracle computes from it
    cam_o = oracle.camera(np.zeros((1, 9), np.float32))
    cam_o.origin[:] = origin.astype(np.float32); cam_o.detector[:] = detector.astype(np.float32)
    cam_o.up[:] = up.astype(np.float32); cam_o.right[:] = right.astype(np.float32)
    cam_o.range[:] = [0.0, 0.0, np.float32(sp * W / 2.0)]
    spacing = float(oracle.pixel_spacing(cam_o, W, H))
    assert abs(spacing / sp - 1) < 1e-5
    cam = Camera()
    cam.origin[:] = cam_o.origin[:]; cam.detector[:] = cam_o.detector[:]; cam.up[:] = cam_o.up[:]; cam.right[:] = cam_o.right[:]
    cam.pixel_spacing[0] = spacing; cam.pixel_spacing[1] = spacing
    cam.width = W; cam.height = H
    small = W * H <= 200 * 200
    tri = _random_soup((origin, detector, up, right, sp, W, H), rng, 400 if small else 250, scale)
    if small:
        ranges = [(0, W * H)]
    else:                                      # 2 rows at the top edge, 3 through the middle, 2 at the bottom edge
        ranges = [(0, 2 * W), ((H // 2 - 1) * W, (H // 2 + 2) * W), ((H - 2) * W, H * W)]
    hits = 0
    for box in ALL_BOXES if small else [(16, 4), (32, 2), (8, 8), (32, 16), (64, 16)]:
        pk = srtlib.debug_prepack(tri, cam, box=box)
        assert pk["unfilterable"] == 0
        for pr in ranges:
            a_ = oracle.filter_audit(tri, cam_o, W, H, pk, pixel_range=pr)
            assert a_["violations"] == 0 and a_["box3_violations"] == 0 and a_["box_violations"] == 0, (frame, scale, box, pr, a_)
            hits += a_["hits"]
    assert hits > 2000, "camera misses the soup: the test would be vacuous"


def test_unfilterable_triangles_always_survive(srtlib, oracle, dragon):
    cam_o = oracle.camera(dragon)
    cam = srtlib.camera_from_bbox(cam_o.vec("lower"), cam_o.vec("upper"), 16, 16)
    tri = dragon[:6].copy()
    tri[1, 4] = np.inf; tri[3, 0] = np.nan; tri[4, :] = 3e38
    pk = srtlib.debug_prepack(tri, cam)
    assert pk["unfilterable"] == 3
    for k in (1, 3, 4):
        assert (pk["filt"][k] == 0).all() and (pk["filt2"][k] == 0).all()


def test_host_copy_pool_moves_every_byte(srtlib):
    """The thread pool behind srt_set_host_copy_threads (csrc/host_copy.h; frames returned into pageable buffers):
    every byte arrives, nothing outside the destination is touched, for any chunk size and thread count."""
    lib = srtlib.lib
    rng = np.random.default_rng(5)
    src = rng.integers(0, 255, 3_000_017, dtype=np.uint8)
    for it in range(200):
        n = int(rng.integers(0, src.size)) if it % 4 else int(rng.integers(0, 5000))
        chunk = int(rng.integers(1, 400_000)) if it % 4 else int(rng.integers(1, 64))
        threads = int(rng.integers(0, 9))
        spin_us = int(rng.choice([0, 1, 50, 2000]))
        dst = np.full(n + 16, 7, np.uint8)
        assert lib.srt_debug_host_copy(dst.ctypes.data + 8, src.ctypes.data, n, chunk, threads, spin_us) == 0
        assert (dst[8:8 + n] == src[:n]).all() and (dst[:8] == 7).all() and (dst[8 + n:] == 7).all()
    assert lib.srt_debug_host_copy(None, None, 0, 1, 0, 0) == 0
    assert lib.srt_debug_host_copy(dst.ctypes.data, src.ctypes.data, 8, 0, 1, 0) != 0       # chunk size 0
    assert lib.srt_debug_host_copy(dst.ctypes.data, src.ctypes.data, 8, 8, 33, 0) != 0      # too many threads
