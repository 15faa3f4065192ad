"""ctypes binding of include/srt.h (tests and bench only; no computation here)."""
import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "lib" / "libsrt_b200.so"

KERNEL_AUTO, KERNEL_EXACT, KERNEL_FILTERED, KERNEL_AUDIT, KERNEL_CULLED = 0, 1, 2, 3, 4
FLAG_ALL_MESHES_CONTRIBUTE = 1
FLAG_FRAME0_MULTICAST = 2
FLAG_L_HALO_ONLY = 4


class SrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"srt error {code}: {msg}")
        self.code = code


class Material(C.Structure):
    _fields_ = [("ambient", C.c_float * 3), ("diffuse", C.c_float * 3), ("specular", C.c_float * 3),
                ("shininess", C.c_float)]


class MeshDesc(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("normals", C.c_void_p), ("texcoords", C.c_void_p),
                ("n_triangles", C.c_uint32), ("mu", C.c_float), ("material", Material),
                ("texture_rgb", C.c_void_p), ("tex_w", C.c_uint32), ("tex_h", C.c_uint32)]


class Camera(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("detector", C.c_float * 3), ("up", C.c_float * 3),
                ("right", C.c_float * 3), ("pixel_spacing", C.c_float * 2),
                ("width", C.c_uint32), ("height", C.c_uint32)]

    def vec(self, name):
        return np.array(getattr(self, name)[:], dtype=np.float32)


class Light(C.Structure):
    _fields_ = [("colour", C.c_float * 3), ("direction", C.c_float * 3), ("position", C.c_float * 3)]


class Params(C.Structure):
    _fields_ = [("incident", C.c_float), ("length_to_cm", C.c_double), ("repair_depth", C.c_uint32),
                ("kernel", C.c_uint32), ("flags", C.c_uint32), ("pixels_per_thread", C.c_uint32)]


class Timing(C.Structure):
    _fields_ = [("render_ms", C.c_float), ("repair_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float),
                ("pair_tests", C.c_uint64), ("exact_tests", C.c_uint64), ("filter_violations", C.c_uint64),
                ("kernel_used", C.c_uint32), ("launches", C.c_uint32), ("main_kernel_ms", C.c_float),
                ("scan_cols", C.c_uint32), ("min_strip_slack", C.c_float), ("min_box_slack", C.c_float),
                ("shadow_grid_ms", C.c_float), ("shadow_grid_cells", C.c_uint32), ("shadow_grid_entries", C.c_uint32),
                ("shadow_grid_rebuilt", C.c_uint32)]


# every symbol include/srt.h declares (tests/test_abi.py checks the .so exports all of them)
EXPORTS = [
    "srt_create", "srt_destroy", "srt_last_error", "srt_version", "srt_set_stream", "srt_synchronize",
    "srt_default_params", "srt_scene_bbox", "srt_camera_from_bbox", "srt_partition_rows", "srt_triangle_normals",
    "srt_set_meshes", "srt_set_camera", "srt_render_lbuffer", "srt_render_lbuffer_device", "srt_repair",
    "srt_repair_device", "srt_render_xray", "srt_interleave_slot_rows",
    "srt_render_lbuffer_interleaved_device", "srt_deinterleave_device", "srt_render_lbuffer_blocks_device", "srt_render_shaded",
    "srt_render_shaded_device", "srt_debug_shaded_buffers", "srt_last_timing", "srt_debug_count_exact", "srt_debug_force_scan_cols", "srt_debug_force_list_caps", "srt_debug_force_shadow_grid", "srt_debug_prepack", "srt_debug_ffma_peak", "srt_debug_pack_info", "srt_debug_download_packed",
    "srt_host_alloc", "srt_host_free", "srt_write_jpeg", "srt_render_xray_sorted", "srt_render_xray_sorted_device",
    "srt_multi_create", "srt_multi_destroy", "srt_multi_last_error", "srt_multi_device_count", "srt_multi_set_meshes",
    "srt_multi_set_camera", "srt_multi_render_xray_device", "srt_multi_synchronize", "srt_multi_frame",
    "srt_multi_render_xray", "srt_multi_last_timing",
    "srt_set_profiling", "srt_set_scan_residency", "srt_peer_barrier_device", "srt_render_xray_blocks_device",
    "srt_repair_blocks_device", "srt_set_host_copy_threads", "srt_debug_host_copy", "srt_debug_host_copy_stats",
    "srt_debug_tune_host_copy",
]


def load_library(path=None):
    path = Path(path) if path else LIB_PATH
    if not path.exists():
        raise FileNotFoundError(
            f"{path} is missing: build it with `make -C simpleraytracing_b200/csrc` "
            "(or __graft_entry__.build()). There is no CPU fallback.")
    return C.CDLL(str(path))


def _ptr(a, dtype):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous, (a.dtype, dtype)
    return a.ctypes.data_as(C.c_void_p)


class Library:
    """Function table; host-only helpers are usable without a GPU."""

    def __init__(self, path=None):
        self.lib = L = load_library(path)
        vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
        L.srt_last_error.restype = C.c_char_p
        L.srt_last_error.argtypes = [vp]
        L.srt_create.argtypes = [C.POINTER(vp), i32]
        L.srt_destroy.argtypes = [vp]
        L.srt_destroy.restype = None
        L.srt_set_stream.argtypes = [vp, vp]
        L.srt_synchronize.argtypes = [vp]
        L.srt_default_params.argtypes = [C.POINTER(Params)]
        L.srt_default_params.restype = None
        L.srt_scene_bbox.argtypes = [C.POINTER(MeshDesc), i32, vp, vp]
        L.srt_camera_from_bbox.argtypes = [vp, vp, u32, u32, C.POINTER(Camera)]
        L.srt_partition_rows.argtypes = [u32, u32, u32, C.POINTER(u32), C.POINTER(u32)]
        L.srt_triangle_normals.argtypes = [vp, u32, vp]
        L.srt_set_meshes.argtypes = [vp, i32, C.POINTER(MeshDesc)]
        L.srt_set_camera.argtypes = [vp, C.POINTER(Camera)]
        L.srt_render_lbuffer.argtypes = [vp, C.POINTER(Params), u32, u32, vp, vp, vp, vp]
        L.srt_render_lbuffer_device.argtypes = [vp, C.POINTER(Params), u32, u32, vp, vp, vp, vp]
        L.srt_interleave_slot_rows.argtypes = [u32, u32, C.POINTER(u32)]
        L.srt_render_lbuffer_interleaved_device.argtypes = [vp, C.POINTER(Params), u32, u32, vp, vp, vp, vp]
        L.srt_render_lbuffer_blocks_device.argtypes = [vp, C.POINTER(Params), u32, u32, u32, C.POINTER(C.c_void_p)]
        L.srt_deinterleave_device.argtypes = [vp, u32, vp, vp]
        L.srt_render_shaded.argtypes = [vp, C.POINTER(Params), C.POINTER(Light), vp, vp]
        L.srt_render_shaded_device.argtypes = [vp, C.POINTER(Params), C.POINTER(Light), vp, vp, vp]
        L.srt_debug_shaded_buffers.argtypes = [vp, vp, vp, vp]
        L.srt_debug_force_shadow_grid.argtypes = [vp, C.c_int]
        L.srt_repair.argtypes = [vp, C.POINTER(Params), vp, vp]
        L.srt_repair_device.argtypes = [vp, C.POINTER(Params), vp, vp]
        L.srt_render_xray.argtypes = [vp, C.POINTER(Params), vp]
        L.srt_last_timing.argtypes = [vp, C.POINTER(Timing)]
        L.srt_debug_count_exact.argtypes = [vp, i32]
        L.srt_debug_force_scan_cols.argtypes = [vp, i32]
        L.srt_debug_force_list_caps.argtypes = [vp, u32, u32, u32]
        L.srt_debug_download_packed.argtypes = [vp, i32, vp, vp, vp]
        L.srt_debug_pack_info.argtypes = [C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
        L.srt_debug_ffma_peak.argtypes = [vp, i32, C.POINTER(C.c_float)]
        L.srt_debug_prepack.argtypes = [vp, u32, C.POINTER(Camera), u32, u32, vp, vp, vp, C.POINTER(u32), C.POINTER(u32)]
        L.srt_render_xray_sorted.argtypes = [vp, C.POINTER(Params), vp]
        L.srt_render_xray_sorted_device.argtypes = [vp, C.POINTER(Params), vp, vp, vp]
        L.srt_write_jpeg.argtypes = [vp, C.c_char_p, vp, u32, u32, i32]
        L.srt_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
        L.srt_host_free.argtypes = [vp]
        L.srt_host_free.restype = None
        L.srt_multi_create.argtypes = [C.POINTER(vp), C.POINTER(i32), i32]
        L.srt_multi_destroy.argtypes = [vp]
        L.srt_multi_destroy.restype = None
        L.srt_multi_last_error.argtypes = [vp]
        L.srt_multi_last_error.restype = C.c_char_p
        L.srt_multi_device_count.argtypes = [vp]
        L.srt_multi_set_meshes.argtypes = [vp, i32, C.POINTER(MeshDesc)]
        L.srt_multi_set_camera.argtypes = [vp, C.POINTER(Camera)]
        L.srt_multi_render_xray_device.argtypes = [vp, C.POINTER(Params)]
        L.srt_multi_synchronize.argtypes = [vp]
        L.srt_multi_frame.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp)]
        L.srt_multi_render_xray.argtypes = [vp, C.POINTER(Params), vp]
        L.srt_multi_last_timing.argtypes = [vp, vp, C.POINTER(C.c_float)]
        L.srt_set_profiling.argtypes = [vp, i32]
        L.srt_set_scan_residency.argtypes = [vp, u32]
        L.srt_set_host_copy_threads.argtypes = [vp, i32]
        L.srt_debug_host_copy.argtypes = [vp, vp, C.c_size_t, C.c_size_t, i32, i32]
        L.srt_debug_host_copy_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.srt_debug_tune_host_copy.argtypes = [vp, C.c_size_t, i32]
        L.srt_peer_barrier_device.argtypes = [vp, u32, u32, C.POINTER(C.c_void_p), u32, u32]
        L.srt_render_xray_blocks_device.argtypes = [vp, C.POINTER(Params), u32, u32, u32, C.POINTER(C.c_void_p),
                                                    C.POINTER(C.c_void_p), u32, vp, u32, C.POINTER(C.c_void_p)]
        L.srt_repair_blocks_device.argtypes = [vp, C.POINTER(Params), u32, u32, vp, u32, C.POINTER(C.c_void_p)]

    # ---- host-only helpers ---------------------------------------------------
    def default_params(self, **kw):
        p = Params()
        self.lib.srt_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    @staticmethod
    def mesh_descs(meshes):
        """meshes: list of (tri[n,9] float32, mu) or of dicts(tri, mu, texcoords, material=(Ka, Kd, Ks, shininess),
        texture uint8[h, w, 3]). Returns (ctypes array, keepalive)."""
        arr = (MeshDesc * len(meshes))()
        keep = []
        for i, m in enumerate(meshes):
            if not isinstance(m, dict):
                m = dict(tri=m[0], mu=m[1])
            tri = np.ascontiguousarray(m["tri"], np.float32).reshape(-1, 9)
            keep.append(tri)
            arr[i].vertices = tri.ctypes.data if len(tri) else None
            arr[i].n_triangles = len(tri)
            arr[i].mu = m.get("mu", 0.0)
            if m.get("texcoords") is not None:
                tc = np.ascontiguousarray(m["texcoords"], np.float32).reshape(-1, 9); keep.append(tc)
                arr[i].texcoords = tc.ctypes.data
            ka, kd, ks, sh = m.get("material", ((1, 1, 1), (1, 1, 1), (1, 1, 1), 1.0))
            arr[i].material.ambient[:] = ka; arr[i].material.diffuse[:] = kd; arr[i].material.specular[:] = ks
            arr[i].material.shininess = sh
            if m.get("texture") is not None:
                t = np.ascontiguousarray(m["texture"], np.uint8); keep.append(t)
                arr[i].texture_rgb = t.ctypes.data; arr[i].tex_h, arr[i].tex_w = t.shape[0], t.shape[1]
        return arr, keep

    def scene_bbox(self, meshes):
        arr, keep = self.mesh_descs(meshes)
        lo = np.empty(3, np.float32); hi = np.empty(3, np.float32)
        rc = self.lib.srt_scene_bbox(arr, len(meshes), _ptr(lo, np.float32), _ptr(hi, np.float32))
        if rc:
            raise SrtError(rc, "srt_scene_bbox")
        return lo, hi

    def camera_from_bbox(self, lo, hi, W, H):
        cam = Camera()
        lo = np.ascontiguousarray(lo, np.float32); hi = np.ascontiguousarray(hi, np.float32)
        rc = self.lib.srt_camera_from_bbox(_ptr(lo, np.float32), _ptr(hi, np.float32), W, H, C.byref(cam))
        if rc:
            raise SrtError(rc, "srt_camera_from_bbox")
        return cam

    def partition_rows(self, H, n, i):
        b = C.c_uint32(); e = C.c_uint32()
        rc = self.lib.srt_partition_rows(H, n, i, C.byref(b), C.byref(e))
        if rc:
            raise SrtError(rc, "srt_partition_rows")
        return b.value, e.value

    def interleave_slot_rows(self, H, n_parts):
        v = C.c_uint32()
        rc = self.lib.srt_interleave_slot_rows(H, n_parts, C.byref(v))
        if rc:
            raise SrtError(rc, "srt_interleave_slot_rows")
        return v.value

    def triangle_normals(self, tri):
        tri = np.ascontiguousarray(tri, np.float32).reshape(-1, 9)
        out = np.empty((len(tri), 3), np.float32)
        rc = self.lib.srt_triangle_normals(_ptr(tri, np.float32), len(tri), _ptr(out, np.float32))
        if rc:
            raise SrtError(rc, "srt_triangle_normals")
        return out

    def pack_info(self):
        """-> (padding granularity in triangles, default box width, default box height) of the filtered kernel"""
        p, w, h = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self.lib.srt_debug_pack_info(C.byref(p), C.byref(w), C.byref(h))
        return p.value, w.value, h.value

    def debug_prepack(self, tri, cam, box=None):
        """Host pre-pack of one mesh; `box` = (width, height) of the filtered kernel's box (None = default)."""
        tri = np.ascontiguousarray(tri, np.float32).reshape(-1, 9)
        n = len(tri)
        pad, bw, bh = self.pack_info()
        if box is not None:
            bw, bh = box
        npad = (n + pad - 1) // pad * pad
        exact = np.zeros((max(n, 1), 16), np.float32)
        filt = np.zeros((max(npad, 1), 12), np.float32)
        filt2 = np.zeros((max(npad, 1), 8), np.float32)
        a = C.c_uint32(); b = C.c_uint32()
        rc = self.lib.srt_debug_prepack(_ptr(tri, np.float32), n, C.byref(cam), bw, bh, _ptr(exact, np.float32),
                                        _ptr(filt, np.float32), _ptr(filt2, np.float32), C.byref(a), C.byref(b))
        if rc:
            raise SrtError(rc, "srt_debug_prepack")
        assert a.value == npad
        return dict(exact=exact[:n], filt=filt[:npad], filt2=filt2[:npad], n_padded=a.value, unfilterable=b.value,
                    box=(bw, bh))


def box_shape(px, py=1, k=0):
    """srt_params.pixels_per_thread encoding of a shape of the filtered kernel: box of px x py pixels, k triangles per
    thread (0 = default)."""
    return px | (py << 8) | (k << 16)


class Context:
    """One GPU context (srt_ctx). Raises SrtError when no sm_100 device is usable."""

    def __init__(self, device=0, library=None):
        self.L = library or Library()
        self.lib = self.L.lib
        h = C.c_void_p()
        rc = self.lib.srt_create(C.byref(h), device)
        if rc:
            raise SrtError(rc, self.lib.srt_last_error(None).decode())
        self.h = h
        self._keep = None
        self.cam = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.srt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise SrtError(rc, self.lib.srt_last_error(self.h).decode())

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.srt_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._ck(self.lib.srt_synchronize(self.h))

    def set_meshes(self, meshes):
        arr, keep = Library.mesh_descs(meshes)
        self._ck(self.lib.srt_set_meshes(self.h, len(meshes), arr))

    def set_camera(self, cam):
        self.cam = cam
        self._ck(self.lib.srt_set_camera(self.h, C.byref(cam)))

    def force_scan_cols(self, cols):
        self._ck(self.lib.srt_debug_force_scan_cols(self.h, int(cols)))

    def force_list_caps(self, bin_cap=0, cand_cap=0, pix_cap=0):
        self._ck(self.lib.srt_debug_force_list_caps(self.h, bin_cap, cand_cap, pix_cap))

    def count_exact(self, on=True):
        self._ck(self.lib.srt_debug_count_exact(self.h, int(on)))

    def render_lbuffer(self, params, rows=None, extras=True, out=None):
        """Host-buffer render of a row band -> dict(L[, dist, ssum, hits]) as full-frame arrays."""
        W, H = self.cam.width, self.cam.height
        rb, re = (0, H) if rows is None else rows
        n = W * H
        out = out or {}
        L = out.get("L"); L = np.zeros(n, np.float32) if L is None else L
        dist = ssum = hits = None
        if extras:
            dist = np.zeros(n, np.float32); ssum = np.zeros(n, np.int32); hits = np.zeros(n, np.uint32)
        self._ck(self.lib.srt_render_lbuffer(self.h, C.byref(params), rb, re, _ptr(L, np.float32),
                                             _ptr(dist, np.float32), _ptr(ssum, np.int32), _ptr(hits, np.uint32)))
        return dict(L=L, dist=dist, ssum=ssum, hits=hits)

    def render_lbuffer_device(self, params, rows, dL, d_dist=0, d_ssum=0, d_hits=0):
        """Device-pointer render (ints = raw device addresses, e.g. torch.Tensor.data_ptr())."""
        self._ck(self.lib.srt_render_lbuffer_device(self.h, C.byref(params), rows[0], rows[1], C.c_void_p(dL),
                                                    C.c_void_p(d_dist or None), C.c_void_p(d_ssum or None),
                                                    C.c_void_p(d_hits or None)))

    def render_lbuffer_blocks_device(self, params, n_parts, part, frame_ptrs, multicast=False, halo=False):
        """Fused render + assembly: this rank's 16-row blocks, stored at image positions into every frame in
        `frame_ptrs` (own frame first, then the peers' NVLink mappings) -- or, with multicast=True, into the single
        NVSwitch multicast mapping frame_ptrs[0]."""
        arr = (C.c_void_p * len(frame_ptrs))(*[C.c_void_p(int(q)) for q in frame_ptrs])
        if multicast or halo:
            q = Params.from_buffer_copy(params)
            q.flags |= (FLAG_FRAME0_MULTICAST if multicast else 0) | (FLAG_L_HALO_ONLY if halo else 0)
            params = q
        self._ck(self.lib.srt_render_lbuffer_blocks_device(self.h, C.byref(params), n_parts, part, len(frame_ptrs), arr))

    def set_profiling(self, on=True):
        self._ck(self.lib.srt_set_profiling(self.h, int(on)))

    def set_scan_residency(self, max_ctas_per_sm=0):
        self._ck(self.lib.srt_set_scan_residency(self.h, int(max_ctas_per_sm)))

    def set_host_copy_threads(self, threads=-1):
        """Helper threads for frames returned into pageable host buffers (-1 = default, 0 = the driver's own copy)."""
        self._ck(self.lib.srt_set_host_copy_threads(self.h, int(threads)))

    def tune_host_copy(self, chunk_bytes=0, spin_us=-1):
        self._ck(self.lib.srt_debug_tune_host_copy(self.h, int(chunk_bytes), int(spin_us)))

    def staged_copies(self):
        n = C.c_uint64(0)
        self._ck(self.lib.srt_debug_host_copy_stats(self.h, C.byref(n)))
        return int(n.value)

    def peer_barrier_device(self, n_ranks, rank, flag_ptrs, epoch, phases=3):
        arr = (C.c_void_p * len(flag_ptrs))(*[C.c_void_p(int(q)) for q in flag_ptrs])
        self._ck(self.lib.srt_peer_barrier_device(self.h, n_ranks, rank, arr, epoch, phases))

    @staticmethod
    def frame_call_args(params, frame_ptrs, flag_ptrs, image_ptrs, multicast=False, halo=False):
        """ctypes arguments of render_xray_blocks_device that do not change from frame to frame (build once per
        frame set: the per-frame call is then one foreign call).  frame_ptrs / image_ptrs: own buffer first, then the
        peers' mappings -- or the single multicast mapping with multicast=True."""
        fr = (C.c_void_p * len(frame_ptrs))(*[C.c_void_p(int(q)) for q in frame_ptrs])
        fl = (C.c_void_p * len(flag_ptrs))(*[C.c_void_p(int(q)) for q in flag_ptrs])
        im = (C.c_void_p * len(image_ptrs))(*[C.c_void_p(int(q)) for q in image_ptrs])
        q = Params.from_buffer_copy(params)
        if multicast:
            q.flags |= FLAG_FRAME0_MULTICAST
        if halo:
            q.flags |= FLAG_L_HALO_ONLY
        return q, fr, len(frame_ptrs), fl, im, len(image_ptrs)

    def render_xray_blocks_device(self, call_args, n_parts, part, epoch, d_own_frame):
        """One frame of the sharded pipeline: this rank's 16-row blocks stored into every rank's L frame, device barrier,
        repair of this rank's blocks (from the assembled frame d_own_frame) stored into every rank's image, device
        barrier.  call_args: frame_call_args(...); epoch = 1, 2, ... per set of flag arrays."""
        q, fr, nf, fl, im, ni = call_args
        self._ck(self.lib.srt_render_xray_blocks_device(self.h, C.byref(q), n_parts, part, nf, fr, fl, epoch,
                                                        C.c_void_p(d_own_frame), ni, im))

    def repair_blocks_device(self, params, n_parts, part, d_L, image_ptrs, multicast=False):
        im = (C.c_void_p * len(image_ptrs))(*[C.c_void_p(int(q)) for q in image_ptrs])
        if multicast:
            q = Params.from_buffer_copy(params)
            q.flags |= FLAG_FRAME0_MULTICAST
            params = q
        self._ck(self.lib.srt_repair_blocks_device(self.h, C.byref(params), n_parts, part, C.c_void_p(d_L), len(image_ptrs), im))

    def render_lbuffer_interleaved_device(self, params, n_parts, part, d_slot, d_dist=0, d_ssum=0, d_hits=0):
        self._ck(self.lib.srt_render_lbuffer_interleaved_device(self.h, C.byref(params), n_parts, part, C.c_void_p(d_slot),
                                                                C.c_void_p(d_dist or None), C.c_void_p(d_ssum or None),
                                                                C.c_void_p(d_hits or None)))

    def deinterleave_device(self, n_parts, d_slots, d_frame):
        self._ck(self.lib.srt_deinterleave_device(self.h, n_parts, C.c_void_p(d_slots), C.c_void_p(d_frame)))

    def render_shaded(self, params, light_position, background=(128, 128, 128), light_colour=(1.0, 1.0, 1.0)):
        """-> uint8 [H*W*3]"""
        n = self.cam.width * self.cam.height
        rgb = np.empty(n * 3, np.uint8)
        lt = Light(); lt.colour[:] = light_colour; lt.position[:] = [float(v) for v in light_position]
        bg = np.array(background, np.uint8)
        self._ck(self.lib.srt_render_shaded(self.h, C.byref(params), C.byref(lt), _ptr(bg, np.uint8), _ptr(rgb, np.uint8)))
        return rgb

    def shaded_buffers(self):
        n = self.cam.width * self.cam.height
        z = np.empty(n, np.float32); tid = np.empty(n, np.int32)
        self._ck(self.lib.srt_debug_shaded_buffers(self.h, _ptr(z, np.float32), _ptr(tid, np.int32), None))
        return z, tid

    def shadow_flags(self):
        """-> uint8 [H*W]: 1 where the shadow ray of the last shaded render was blocked"""
        sh = np.empty(self.cam.width * self.cam.height, np.uint8)
        self._ck(self.lib.srt_debug_shaded_buffers(self.h, None, None, _ptr(sh, np.uint8)))
        return sh

    def force_shadow_grid(self, cells_per_side):
        self._ck(self.lib.srt_debug_force_shadow_grid(self.h, int(cells_per_side)))

    def repair(self, params, L):
        L = np.ascontiguousarray(L, np.float32).reshape(-1)
        img = np.empty_like(L)
        self._ck(self.lib.srt_repair(self.h, C.byref(params), _ptr(L, np.float32), _ptr(img, np.float32)))
        return img

    def repair_device(self, params, dL, d_img):
        self._ck(self.lib.srt_repair_device(self.h, C.byref(params), C.c_void_p(dL), C.c_void_p(d_img)))

    def render_xray(self, params, out=None):
        n = self.cam.width * self.cam.height
        img = np.empty(n, np.float32) if out is None else out
        self._ck(self.lib.srt_render_xray(self.h, C.byref(params), _ptr(img, np.float32)))
        return img

    def render_xray_sorted(self, params, out=None):
        """Sort-based X-ray image of src/main.cxx (no repair pass) -> float32 [H*W]"""
        n = self.cam.width * self.cam.height
        img = np.empty(n, np.float32) if out is None else out
        self._ck(self.lib.srt_render_xray_sorted(self.h, C.byref(params), _ptr(img, np.float32)))
        return img

    def render_xray_sorted_device(self, params, d_image, d_path=0, d_hits=0):
        self._ck(self.lib.srt_render_xray_sorted_device(self.h, C.byref(params), C.c_void_p(d_image), C.c_void_p(d_path or None),
                                                        C.c_void_p(d_hits or None)))

    def write_jpeg(self, path, rgb, quality=100):
        """rgb: uint8 [H, W, 3] -> JPEG file (nvJPEG encoder on this context's GPU)"""
        rgb = np.ascontiguousarray(rgb, np.uint8)
        self._ck(self.lib.srt_write_jpeg(self.h, str(path).encode(), _ptr(rgb, np.uint8), rgb.shape[1], rgb.shape[0], quality))

    def download_packed(self, mesh, n):
        pad = self.L.pack_info()[0]
        npad = (n + pad - 1) // pad * pad
        exact = np.zeros((max(n, 1), 16), np.float32); filt = np.zeros((max(npad, 1), 12), np.float32)
        filt2 = np.zeros((max(npad, 1), 8), np.float32)
        self._ck(self.lib.srt_debug_download_packed(self.h, mesh, _ptr(exact, np.float32), _ptr(filt, np.float32),
                                                    _ptr(filt2, np.float32)))
        return dict(exact=exact[:n], filt=filt[:npad], filt2=filt2[:npad], n_padded=npad)

    def ffma_peak_tflops(self, repeats=3):
        v = C.c_float()
        self._ck(self.lib.srt_debug_ffma_peak(self.h, repeats, C.byref(v)))
        return float(v.value)

    def timing(self):
        t = Timing()
        self._ck(self.lib.srt_last_timing(self.h, C.byref(t)))
        return t


class PinnedArray:
    """numpy view of page-locked host memory from srt_host_alloc (freed with the object)."""

    def __init__(self, library, shape, dtype):
        self.lib = library.lib
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        rc = self.lib.srt_host_alloc(C.byref(p), self.nbytes)
        if rc:
            raise SrtError(rc, self.lib.srt_last_error(None).decode())
        self.ptr = p
        buf = (C.c_char * self.nbytes).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def close(self):
        if getattr(self, "ptr", None):
            self.array = None
            self.lib.srt_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiContext:
    """srt_multi: one process driving N GPUs (render kernels store into every GPU's frame over NVLink peer mappings)."""

    def __init__(self, devices, library=None):
        self.L = library or Library()
        self.lib = self.L.lib
        ids = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.lib.srt_multi_create(C.byref(h), ids, len(devices))
        if rc:
            raise SrtError(rc, self.lib.srt_multi_last_error(None).decode())
        self.h = h
        self.n = len(devices)
        self.cam = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.srt_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise SrtError(rc, self.lib.srt_multi_last_error(self.h).decode())

    def set_meshes(self, meshes):
        arr, keep = Library.mesh_descs(meshes)
        self._ck(self.lib.srt_multi_set_meshes(self.h, len(meshes), arr))

    def set_camera(self, cam):
        self.cam = cam
        self._ck(self.lib.srt_multi_set_camera(self.h, C.byref(cam)))

    def render_xray(self, params, out=None):
        n = self.cam.width * self.cam.height
        img = np.empty(n, np.float32) if out is None else out
        self._ck(self.lib.srt_multi_render_xray(self.h, C.byref(params), _ptr(img, np.float32)))
        return img

    def render_xray_device(self, params):
        self._ck(self.lib.srt_multi_render_xray_device(self.h, C.byref(params)))

    def synchronize(self):
        self._ck(self.lib.srt_multi_synchronize(self.h))

    def frame(self, index):
        """-> (device pointer of the pre-repair L-buffer, device pointer of the repaired image) on device `index`"""
        a = C.c_void_p(); b = C.c_void_p()
        self._ck(self.lib.srt_multi_frame(self.h, index, C.byref(a), C.byref(b)))
        return a.value, b.value

    def timing(self):
        """-> (render kernel ms per device, step ms of the slowest device)"""
        per = np.zeros(self.n, np.float32); step = C.c_float()
        self._ck(self.lib.srt_multi_last_timing(self.h, _ptr(per, np.float32), C.byref(step)))
        return per, float(step.value)
