"""ctypes bindings for the parity oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Nothing under simpleraytracing_b200/ does.

  Oracle      oracle/liboracle.so       plain-C restatement (srt_oracle.c)
  Reference   oracle/_ref/libsrt_ref.so the reference's own Ray/Triangle/TriangleMesh
                                        sources behind the same interface (ref_shim.cpp)
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


class Camera(C.Structure):
    """RayTracerInfo (main-pthreads-lbuffer.cxx:125-135) as plain floats."""
    _fields_ = [(n, C.c_float * 3) for n in
                ("detector", "origin", "up", "right", "lower", "upper", "range",
                 "light_position", "light_direction")]

    def vec(self, name):
        return np.array(getattr(self, name)[:], dtype=np.float32)


class Material(C.Structure):
    _fields_ = [("ambient", C.c_float * 3), ("diffuse", C.c_float * 3), ("specular", C.c_float * 3),
                ("shininess", C.c_float)]


def _shaded_args(meshes):
    """meshes: list of dict(tri[n,9], texcoords[n,9] or None, material=(Ka, Kd, Ks, shininess), texture uint8[h,w,3] or None)"""
    tri = np.ascontiguousarray(np.concatenate([np.asarray(m["tri"], np.float32).reshape(-1, 9) for m in meshes]))
    have_tc = any(m.get("texcoords") is not None for m in meshes)
    tc = None
    if have_tc:
        tc = np.ascontiguousarray(np.concatenate([
            np.asarray(m["texcoords"], np.float32).reshape(-1, 9) if m.get("texcoords") is not None
            else np.zeros((len(m["tri"]), 9), np.float32) for m in meshes]))
    nt = np.array([len(m["tri"]) for m in meshes], np.uint32)
    mats = (Material * len(meshes))()
    texs = (C.c_void_p * len(meshes))()
    tw = np.zeros(len(meshes), np.uint32); th = np.zeros(len(meshes), np.uint32)
    keep = []
    for i, m in enumerate(meshes):
        ka, kd, ks, sh = m["material"]
        mats[i].ambient[:] = ka; mats[i].diffuse[:] = kd; mats[i].specular[:] = ks; mats[i].shininess = sh
        t = m.get("texture")
        if t is not None:
            t = np.ascontiguousarray(t, np.uint8); keep.append(t)
            texs[i] = t.ctypes.data; th[i], tw[i] = t.shape[0], t.shape[1]
    return tri, tc, nt, mats, texs, tw, th, keep


def build(force=False):
    """Compile liboracle.so (and _ref/ when /root/reference is present)."""
    subprocess.run(["make", "-C", str(HERE), str(HERE / "liboracle.so")] + (["-B"] if force else []),
                   check=True, stdout=subprocess.DEVNULL)
    if Path("/root/reference/src").is_dir():
        subprocess.run(["make", "-C", str(HERE), "ref"], check=True, stdout=subprocess.DEVNULL)


def _opt(a, dtype):
    if a is None:
        return None
    assert a.dtype == dtype and a.flags.c_contiguous
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Plain-C restatement."""

    def __init__(self):
        build()
        self.lib = C.CDLL(str(HERE / "liboracle.so"))
        L = self.lib
        L.oracle_bbox.argtypes = [_f32p, C.c_uint64, _f32p, _f32p]
        L.oracle_init_camera.argtypes = [_f32p, _f32p, C.POINTER(Camera)]
        L.oracle_pixel_spacing.argtypes = [C.POINTER(Camera), C.c_uint32, C.c_uint32]
        L.oracle_pixel_spacing.restype = C.c_float
        L.oracle_make_ray.argtypes = [C.POINTER(Camera), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _f32p, _f32p]
        L.oracle_intersect.argtypes = [_f32p, _f32p, _f32p, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.oracle_ray_direction.argtypes = [_f32p, _f32p]
        L.oracle_triangle_normal.argtypes = [_f32p, _f32p]
        L.oracle_triangle_area.argtypes = [_f32p]
        L.oracle_triangle_area.restype = C.c_float
        L.oracle_render_lbuffer_pixels.argtypes = [_f32p, _u32p, C.c_int, C.POINTER(Camera), C.c_uint32, C.c_uint32,
                                                   C.c_void_p, C.c_uint64, C.c_int, _f32p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_render_lbuffer.argtypes = [_f32p, _u32p, C.c_int, C.POINTER(Camera), C.c_uint32, C.c_uint32,
                                            C.c_uint64, C.c_uint64, C.c_int, _f32p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p]
        L.oracle_repair.argtypes = [_f32p, C.c_uint32, C.c_uint32, _f32p]
        L.oracle_render_sort.argtypes = [_f32p, _u32p, C.c_int, C.POINTER(Camera), C.c_uint32, C.c_uint32,
                                         C.c_uint64, C.c_uint64, _f32p, C.c_void_p]
        L.oracle_save_text.argtypes = [C.c_char_p, _f32p, C.c_uint32, C.c_uint32]
        L.oracle_partition.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
        L.oracle_render_shaded.argtypes = [_f32p, C.c_void_p, _u32p, C.c_int, C.POINTER(Material), C.c_void_p, _u32p, _u32p,
                                           C.POINTER(Camera), C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_filter_audit.argtypes = [_f32p, C.c_uint32, C.POINTER(Camera), C.c_uint32, C.c_uint32, _f32p, _f32p,
                                          C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]

    # scene ---------------------------------------------------------------
    def camera(self, tri):
        tri = np.ascontiguousarray(tri, np.float32).reshape(-1)
        lo = np.empty(3, np.float32); hi = np.empty(3, np.float32)
        self.lib.oracle_bbox(tri, tri.size // 9, lo, hi)
        cam = Camera()
        self.lib.oracle_init_camera(lo, hi, C.byref(cam))
        return cam

    def pixel_spacing(self, cam, W, H):
        return np.float32(self.lib.oracle_pixel_spacing(C.byref(cam), W, H))

    def make_ray(self, cam, W, H, row, col):
        a = np.empty(3, np.float32); b = np.empty(3, np.float32)
        self.lib.oracle_make_ray(C.byref(cam), W, H, row, col, a, b)
        return a, b

    def ray_direction(self, d):
        out = np.empty(3, np.float32)
        self.lib.oracle_ray_direction(np.ascontiguousarray(d, np.float32), out)
        return out

    def intersect(self, o, d, tri9):
        t = C.c_float(0); st = C.c_int(-1)
        hit = self.lib.oracle_intersect(np.ascontiguousarray(o, np.float32), np.ascontiguousarray(d, np.float32),
                                        np.ascontiguousarray(tri9, np.float32).reshape(-1), C.byref(t), C.byref(st))
        return bool(hit), np.float32(t.value), st.value

    def triangle_normal(self, tri9):
        n = np.empty(3, np.float32)
        self.lib.oracle_triangle_normal(np.ascontiguousarray(tri9, np.float32).reshape(-1), n)
        return n

    def triangle_area(self, tri9):
        return np.float32(self.lib.oracle_triangle_area(np.ascontiguousarray(tri9, np.float32).reshape(-1)))

    # renders --------------------------------------------------------------
    def render_lbuffer(self, tri, mesh_ntris, cam, W, H, pixel_range=None, threads=None, want_stages=False):
        """-> dict(L, dist, ssum, hits[, stages]); arrays are W*H (only the range is written)."""
        tri = np.ascontiguousarray(tri, np.float32).reshape(-1)
        nt = np.ascontiguousarray(mesh_ntris, np.uint32)
        assert tri.size == 9 * int(nt.sum())
        n = W * H
        b, e = (0, n - 1) if pixel_range is None else pixel_range
        threads = threads or os.cpu_count() or 1
        out = dict(L=np.zeros(n, np.float32), dist=np.zeros(n, np.float32),
                   ssum=np.zeros(n, np.int32), hits=np.zeros(n, np.uint32))
        stages = np.zeros(4, np.uint64)
        if e >= b:
            self.lib.oracle_render_lbuffer(tri, nt, len(nt), C.byref(cam), W, H, b, e, threads, out["L"],
                                           _opt(out["dist"], np.float32), _opt(out["ssum"], np.int32),
                                           _opt(out["hits"], np.uint32),
                                           stages.ctypes.data_as(C.c_void_p) if want_stages else None)
        if want_stages:
            out["stages"] = stages
        return out

    def render_lbuffer_pixels(self, tri, mesh_ntris, cam, W, H, pixels, threads=None):
        """L-buffer values of a LIST of pixel indices -> dict(L, dist, ssum, hits), one entry per listed pixel."""
        tri = np.ascontiguousarray(tri, np.float32).reshape(-1)
        nt = np.ascontiguousarray(mesh_ntris, np.uint32)
        pix = np.ascontiguousarray(pixels, np.uint64)
        n = W * H
        full = dict(L=np.zeros(n, np.float32), dist=np.zeros(n, np.float32), ssum=np.zeros(n, np.int32), hits=np.zeros(n, np.uint32))
        self.lib.oracle_render_lbuffer_pixels(tri, nt, len(nt), C.byref(cam), W, H, pix.ctypes.data_as(C.c_void_p), len(pix),
                                              threads or os.cpu_count() or 1, full["L"], full["dist"].ctypes.data_as(C.c_void_p),
                                              full["ssum"].ctypes.data_as(C.c_void_p), full["hits"].ctypes.data_as(C.c_void_p))
        idx = pix.astype(np.int64)
        return {k: v[idx] for k, v in full.items()}

    def render_shaded(self, meshes, cam, W, H, background=(128, 128, 128), pixel_range=None, threads=None):
        """-> dict(rgb uint8[H*W*3], z float32, tri_id int32, shadow uint8). cam = camera WITHOUT the background quad."""
        tri, tc, nt, mats, texs, tw, th, keep = _shaded_args(meshes)
        n = W * H
        b, e = (0, n) if pixel_range is None else pixel_range
        rgb = np.tile(np.array(background, np.uint8), n)
        z = np.full(n, np.inf, np.float32); tid = np.full(n, -1, np.int32); sh = np.zeros(n, np.uint8)
        self.lib.oracle_render_shaded(tri.reshape(-1), tc.ctypes.data_as(C.c_void_p) if tc is not None else None, nt, len(nt),
                                      mats, C.cast(texs, C.c_void_p), tw, th, C.byref(cam), W, H, b, e,
                                      threads or os.cpu_count() or 1, rgb.ctypes.data_as(C.c_void_p),
                                      z.ctypes.data_as(C.c_void_p), tid.ctypes.data_as(C.c_void_p), sh.ctypes.data_as(C.c_void_p))
        return dict(rgb=rgb, z=z, tri_id=tid, shadow=sh)

    def repair(self, L, W, H):
        img = np.empty(W * H, np.float32)
        self.lib.oracle_repair(np.ascontiguousarray(L, np.float32).reshape(-1), W, H, img)
        return img

    def render_sort(self, tri, mesh_ntris, cam, W, H, threads=None):
        tri = np.ascontiguousarray(tri, np.float32).reshape(-1)
        nt = np.ascontiguousarray(mesh_ntris, np.uint32)
        n = W * H
        img = np.zeros(n, np.float32); hits = np.zeros(n, np.uint32)
        threads = threads or os.cpu_count() or 1
        cuts = np.linspace(0, n, threads + 1).astype(np.int64)

        def job(i):
            if cuts[i + 1] > cuts[i]:
                self.lib.oracle_render_sort(tri, nt, len(nt), C.byref(cam), W, H, int(cuts[i]), int(cuts[i + 1] - 1),
                                            img, hits.ctypes.data_as(C.c_void_p))
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(job, range(threads)))
        return img, hits

    def save_text(self, path, img, W, H):
        return self.lib.oracle_save_text(str(path).encode(), np.ascontiguousarray(img, np.float32).reshape(-1), W, H)

    def filter_audit(self, tri, cam, W, H, packed, pixel_range=None, threads=None):
        """Audit the product's packed filter constants against the oracle's intersect.
        -> dict(pairs, survive0_box, survive0, survive1, hits, violations, box3_violations, box_violations)"""
        tri = np.ascontiguousarray(tri, np.float32).reshape(-1)
        n = tri.size // 9
        b, e = (0, W * H) if pixel_range is None else pixel_range
        counts = np.zeros(8, np.uint64)
        rc = self.lib.oracle_filter_audit(tri, n, C.byref(cam), W, H,
                                          np.ascontiguousarray(packed["filt"], np.float32).reshape(-1),
                                          np.ascontiguousarray(packed["filt2"], np.float32).reshape(-1),
                                          int(packed["box"][0]), int(packed["box"][1]), b, e,
                                          threads or os.cpu_count() or 1, counts.ctypes.data_as(C.c_void_p))
        if rc:
            raise RuntimeError("host CPU has no FMA: cannot emulate the kernel's FFMA sequence")
        return dict(zip(("pairs", "survive0_box", "survive0", "survive1", "hits", "violations", "box3_violations", "box_violations"),
                        map(int, counts)))

    def partition(self, total, n):
        s = np.zeros(n, np.uint64); e = np.zeros(n, np.uint64)
        self.lib.oracle_partition(total, n, s.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p))
        return s, e


class Reference:
    """The reference's own classes (oracle/_ref/libsrt_ref.so). May be absent."""

    @staticmethod
    def available():
        return (HERE / "_ref" / "libsrt_ref.so").exists()

    def __init__(self):
        build()
        self.lib = C.CDLL(str(HERE / "_ref" / "libsrt_ref.so"))
        L = self.lib
        L.ref_scene_camera.argtypes = [_f32p, _u32p, C.c_int, C.POINTER(Camera)]
        L.ref_intersect.argtypes = [_f32p, _f32p, _f32p, C.POINTER(C.c_float)]
        L.ref_ray_direction.argtypes = [_f32p, _f32p]
        L.ref_triangle_normal.argtypes = [_f32p, _f32p]
        L.ref_triangle_area.argtypes = [_f32p]
        L.ref_triangle_area.restype = C.c_float
        L.ref_render_lbuffer.argtypes = [_f32p, _u32p, C.c_int, C.POINTER(Camera), C.c_uint32, C.c_uint32,
                                         C.c_uint64, C.c_uint64, C.c_int, _f32p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_render_sort.argtypes = [_f32p, _u32p, C.c_int, C.POINTER(Camera), C.c_uint32, C.c_uint32,
                                      C.c_uint64, C.c_uint64, _f32p]
        L.ref_save_text.argtypes = [C.c_char_p, _f32p, C.c_uint32, C.c_uint32]
        L.ref_render_shaded.argtypes = [_f32p, C.c_void_p, _u32p, C.c_int, C.POINTER(Material), C.c_void_p, _u32p, _u32p,
                                        C.POINTER(Camera), C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]

    def render_shaded(self, meshes, cam, W, H, background=(128, 128, 128), threads=None):
        tri, tc, nt, mats, texs, tw, th, keep = _shaded_args(meshes)
        n = W * H
        rgb = np.tile(np.array(background, np.uint8), n)
        threads = threads or os.cpu_count() or 1
        cuts = np.linspace(0, n, threads + 1).astype(np.int64)

        def job(i):
            if cuts[i + 1] > cuts[i]:
                self.lib.ref_render_shaded(tri.reshape(-1), tc.ctypes.data_as(C.c_void_p) if tc is not None else None, nt,
                                           len(nt), mats, C.cast(texs, C.c_void_p), tw, th, C.byref(cam), W, H,
                                           int(cuts[i]), int(cuts[i + 1]), rgb.ctypes.data_as(C.c_void_p))
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(job, range(threads)))
        return rgb

    def camera(self, tri, mesh_ntris):
        tri = np.ascontiguousarray(tri, np.float32).reshape(-1)
        cam = Camera()
        self.lib.ref_scene_camera(tri, np.ascontiguousarray(mesh_ntris, np.uint32), len(mesh_ntris), C.byref(cam))
        return cam

    def intersect(self, o, d, tri9):
        t = C.c_float(0)
        hit = self.lib.ref_intersect(np.ascontiguousarray(o, np.float32), np.ascontiguousarray(d, np.float32),
                                     np.ascontiguousarray(tri9, np.float32).reshape(-1), C.byref(t))
        return bool(hit), np.float32(t.value)

    def ray_direction(self, d):
        out = np.empty(3, np.float32)
        self.lib.ref_ray_direction(np.ascontiguousarray(d, np.float32), out)
        return out

    def triangle_normal(self, tri9):
        n = np.empty(3, np.float32)
        self.lib.ref_triangle_normal(np.ascontiguousarray(tri9, np.float32).reshape(-1), n)
        return n

    def triangle_area(self, tri9):
        return np.float32(self.lib.ref_triangle_area(np.ascontiguousarray(tri9, np.float32).reshape(-1)))

    def render_lbuffer(self, tri, mesh_ntris, cam, W, H, pixel_range=None, threads=None):
        tri = np.ascontiguousarray(tri, np.float32).reshape(-1)
        nt = np.ascontiguousarray(mesh_ntris, np.uint32)
        n = W * H
        b, e = (0, n - 1) if pixel_range is None else pixel_range
        threads = threads or os.cpu_count() or 1
        out = dict(L=np.zeros(n, np.float32), dist=np.zeros(n, np.float32),
                   ssum=np.zeros(n, np.int32), hits=np.zeros(n, np.uint32))
        if e >= b:
            self.lib.ref_render_lbuffer(tri, nt, len(nt), C.byref(cam), W, H, b, e, threads, out["L"],
                                        _opt(out["dist"], np.float32), _opt(out["ssum"], np.int32),
                                        _opt(out["hits"], np.uint32))
        return out

    def render_sort(self, tri, mesh_ntris, cam, W, H, threads=None):
        tri = np.ascontiguousarray(tri, np.float32).reshape(-1)
        nt = np.ascontiguousarray(mesh_ntris, np.uint32)
        n = W * H
        img = np.zeros(n, np.float32)
        threads = threads or os.cpu_count() or 1
        cuts = np.linspace(0, n, threads + 1).astype(np.int64)

        def job(i):
            if cuts[i + 1] > cuts[i]:
                self.lib.ref_render_sort(tri, nt, len(nt), C.byref(cam), W, H, int(cuts[i]), int(cuts[i + 1] - 1), img)
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(job, range(threads)))
        return img

    def save_text(self, path, img, W, H):
        return self.lib.ref_save_text(str(path).encode(), np.ascontiguousarray(img, np.float32).reshape(-1), W, H)
