"""The C++ host side (include/srt/*.h, CLI driver): reference-compatible classes, PLY loader, and -- on a GPU --
the drop-in renderLoop and the main-pthreads-lbuffer-compatible command line."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_golden

BIN = ROOT / "simpleraytracing_b200" / "bin"


@pytest.fixture(scope="module")
def host_bins(srtlib):
    subprocess.run(["make", "-C", str(ROOT / "simpleraytracing_b200" / "host")], check=True, stdout=subprocess.DEVNULL)
    return BIN


def write_ply(path, vertices, faces, binary=True):
    v = np.asarray(vertices, "<f4"); f = np.asarray(faces, "<i4")
    hdr = ("ply\nformat %s 1.0\ncomment test\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
           "element face %d\nproperty list uchar int vertex_indices\nend_header\n") % (
               "binary_little_endian" if binary else "ascii", len(v), len(f))
    with open(path, "wb") as fh:
        fh.write(hdr.encode())
        if binary:
            fh.write(v.tobytes())
            rec = np.zeros(len(f), dtype=[("n", "u1"), ("i", "<i4", 3)])
            rec["n"] = 3; rec["i"] = f
            fh.write(rec.tobytes())
        else:
            for p in v:
                fh.write(("%.9g %.9g %.9g\n" % tuple(p)).encode())
            for t in f:
                fh.write(("3 %d %d %d\n" % tuple(t)).encode())


def expected_sums(vertices, faces):
    soup = np.asarray(vertices, np.float32)[np.asarray(faces).reshape(-1)].reshape(-1).astype(np.float64)
    w = (np.arange(soup.size) % 97) + 1.0
    return len(faces), soup.sum(), (soup * w).sum()


def test_host_classes(host_bins):
    r = subprocess.run([str(host_bins / "test_host_api"), "cpu"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr


def test_host_copy_pool_reused_across_frames(host_bins):
    """csrc/host_copy.h as a context uses it: one pool, many frames of activate -> submit ... -> drain (several rounds),
    for 0 - 7 helper threads and polling times 0 / 20 / 2000 us."""
    r = subprocess.run([str(host_bins / "test_host_api"), "pool"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr


def test_host_classes_equal_the_reference_classes_bit_for_bit(host_bins, tmp_path, oracle, reference, dragon):
    """srt::Ray (constructor + intersect) and srt::Triangle (cached normal, area) of include/srt/*.h against the
    reference's own classes compiled from /root/reference (oracle/_ref) and the C oracle, on dragon triangles under
    camera rays, random soups and the degenerate inputs the reference's code paths distinguish."""
    rng = np.random.default_rng(17)
    recs = []
    cam = oracle.camera(dragon)
    W = H = 256
    for _ in range(1500):                                   # rays of the 256^2 camera against dragon triangles near them
        row, col = int(rng.integers(0, H)), int(rng.integers(0, W))
        _, d = oracle.make_ray(cam, W, H, row, col)
        tri = dragon[int(rng.integers(0, len(dragon)))]
        # aim half of the rays at the triangle so that hits, edge cases and misses all occur
        o = cam.vec("origin")
        if rng.random() < 0.5:
            w = rng.dirichlet([1, 1, 1]).astype(np.float32)
            if rng.random() < 0.3:
                w = np.array([[1, 0, 0], [0, 1, 0], [0.5, 0.5, 0], [0, 0.5, 0.5]][int(rng.integers(0, 4))], np.float32)
            d = (tri.reshape(3, 3) * w[:, None]).sum(0).astype(np.float32) - o
        recs.append(np.concatenate([o, d, tri.reshape(-1)]))
    for _ in range(1500):                                   # random soups over many scales, rays from anywhere
        s = np.float32(10.0 ** rng.uniform(-6, 6))
        tri = (rng.standard_normal(9) * s).astype(np.float32)
        o = (rng.standard_normal(3) * s * 3).astype(np.float32)
        d = (tri.reshape(3, 3).mean(0) + rng.standard_normal(3) * s * 0.3 - o).astype(np.float32)
        if rng.random() < 0.3:
            d = -d                                          # the triangle lies behind the origin: negative t is accepted
        recs.append(np.concatenate([o, d, tri]))
    z = np.zeros(3, np.float32)
    t0 = np.array([0, 0, 0, 2, 0, 0, 0, 2, 0], np.float32)
    recs += [np.concatenate([[0.5, 0.5, -1], z, t0]),                            # zero direction stays zero
             np.concatenate([[0.5, 0.5, -1], [0, 0, 1e-30], t0]),                # tiny direction: length underflows
             np.concatenate([[0.5, 0.5, 0], [1, 0, 0], t0]),                     # ray in the triangle's plane: det == 0
             np.concatenate([[0, 0, -1], [0, 0, 1], t0]),                        # through a vertex
             np.concatenate([[1, 0, -1], [0, 0, 1], t0]),                        # through an edge midpoint
             np.concatenate([[1, 1, -1], [0, 0, 1], t0]),                        # on the hypotenuse: u + v == 1
             np.concatenate([[0.5, 0.5, -1], [0, 0, 1], np.array([0, 0, 0, 1, 1, 1, 2, 2, 2], np.float32)]),   # zero-area triangle
             np.concatenate([[0.5, 0.5, -1], [0, 0, 1], np.zeros(9, np.float32)])]
    recs = np.asarray(recs, np.float32)
    (tmp_path / "in.bin").write_bytes(recs.tobytes())
    r = subprocess.run([str(host_bins / "test_host_api"), "parity", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")],
                       capture_output=True, text=True)
    assert r.returncode == 0 and f"OK {len(recs)} records" in r.stdout, r.stdout + r.stderr
    got = np.frombuffer((tmp_path / "out.bin").read_bytes(), np.float32).reshape(len(recs), 9)

    def same(a, b):                                         # bit for bit; NaN (0 / 0 of a degenerate normal) equals NaN
        a = np.asarray(a, np.float32).reshape(-1); b = np.asarray(b, np.float32).reshape(-1)
        return bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all())

    hits = 0
    for rec, g in zip(recs, got):
        o, d, tri = rec[0:3], rec[3:6], rec[6:15]
        for impl in (reference, oracle):
            assert same(g[0:3], impl.ray_direction(d)), ("direction", rec)
            # the reference's Ray constructor renormalises d itself; the C oracle's intersect takes the stored direction
            res = impl.intersect(o, d if impl is reference else impl.ray_direction(d), tri)
            assert bool(g[3]) == res[0], ("hit flag", rec)
            if res[0]:
                assert same(g[4], res[1]), ("t", rec)
            assert same(g[5:8], impl.triangle_normal(tri)), ("normal", rec)
            assert same(g[8], impl.triangle_area(tri)), ("area", rec)
        hits += int(g[3])
    assert 300 < hits < len(recs) - 300                     # both outcomes are well represented


@pytest.mark.parametrize("binary", [True, False])
def test_ply_loader_matches_fixture(host_bins, tmp_path, binary):
    d = np.load(GOLDEN / "dragon_mesh.npz")
    v, f = d["vertices"], d["faces"]
    if not binary:
        v, f = v[:3000], f[(f < 3000).all(1)][:4000]
    p = tmp_path / "mesh.ply"
    write_ply(p, v, f, binary)
    r = subprocess.run([str(host_bins / "test_host_api"), "ply", str(p)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    n, s, ws = r.stdout.split()
    en, es, ews = expected_sums(v, f)
    assert int(n) == en and float(s) == pytest.approx(es, rel=1e-12) and float(ws) == pytest.approx(ews, rel=1e-12)


def _write_general_ply(path, v, f, fmt, vertex_dtype="f4", index_type=("uchar", "int"), extras=True):
    """A PLY as scanners / modelling tools write them: normals and colours next to x, y, z, a confidence value BEFORE the
    coordinates, per-face flags and texture-coordinate lists, an `edge` element between vertices and faces, quads among the
    triangles; fmt: ascii | binary_little_endian | binary_big_endian."""
    e = {"ascii": "", "binary_little_endian": "<", "binary_big_endian": ">"}[fmt]
    np_of = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4", "double": "f8"}
    vt = "double" if vertex_dtype == "f8" else "float"
    hdr = ["ply", f"format {fmt} 1.0", "comment written by the tests", f"element vertex {len(v)}"]
    if extras:
        hdr += ["property float confidence"]
    hdr += [f"property {vt} x", f"property {vt} y", f"property {vt} z"]
    if extras:
        hdr += ["property float nx", "property float ny", "property float nz", "property uchar red", "property uchar green",
                "property uchar blue", "property short tag", f"element edge 2", "property int vertex1", "property int vertex2",
                "property list uchar ushort marks"]
    hdr += [f"element face {len(f)}"]
    if extras:
        hdr += ["property uchar flags"]
    hdr += [f"property list {index_type[0]} {index_type[1]} vertex_indices"]
    if extras:
        hdr += ["property list uchar float texcoord"]
    hdr += ["end_header"]
    rng = np.random.default_rng(0)
    with open(path, "wb") as fh:
        fh.write(("\n".join(hdr) + "\n").encode())

        def put(kind, value):
            if fmt == "ascii":
                fh.write((("%.17g" if kind in ("float", "double") else "%d") % value + " ").encode())
            else:
                fh.write(np.array(value, dtype=e + np_of[kind]).tobytes())

        for p in v:
            if extras:
                put("float", float(np.float32(rng.random())))
            for c in p:
                put(vt, float(c))
            if extras:
                for c in rng.normal(size=3):
                    put("float", float(np.float32(c)))
                for c in rng.integers(0, 256, 3):
                    put("uchar", int(c))
                put("short", int(rng.integers(-30000, 30000)))
            if fmt == "ascii":
                fh.write(b"\n")
        if extras:
            for k in range(2):
                put("int", k); put("int", k + 1); put("uchar", 3)
                for m in range(3):
                    put("ushort", 1000 * m + k)
                if fmt == "ascii":
                    fh.write(b"\n")
        for t in f:
            if extras:
                put("uchar", int(rng.integers(0, 256)))
            put(index_type[0], len(t))
            for i in t:
                put(index_type[1], int(i))
            if extras:
                put("uchar", 2 * len(t))
                for c in rng.random(2 * len(t)):
                    put("float", float(np.float32(c)))
            if fmt == "ascii":
                fh.write(b"\n")


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
@pytest.mark.parametrize("vertex_dtype,index_type,extras", [("f4", ("uchar", "int"), True), ("f8", ("ushort", "uint"), True),
                                                            ("f4", ("uint", "ushort"), False), ("f4", ("uchar", "uchar"), True)])
def test_ply_loader_general_layouts(host_bins, tmp_path, fmt, vertex_dtype, index_type, extras):
    """The loader stands in for Assimp (lbuffer.cxx:518-582): any PLY encoding, property set and element order; only x, y, z
    and the triangles' vertex_indices are kept, float coordinates bit for bit, in file order."""
    d = np.load(GOLDEN / "dragon_mesh.npz")
    nv = 200 if index_type[1] == "uchar" else 2500
    v = d["vertices"][:nv]
    f = d["faces"]
    f = f[(f < nv).all(1)][:3000]
    faces = [list(t) for t in f]
    faces.insert(5, [0, 1, 2, 3]); faces.insert(40, [4, 5, 6, 7, 8]); faces.append([0, 1])       # skipped: not triangles
    p = tmp_path / "mesh.ply"
    _write_general_ply(p, v.astype(np.float64) if vertex_dtype == "f8" else v, faces, fmt, vertex_dtype, index_type, extras)
    r = subprocess.run([str(host_bins / "test_host_api"), "ply", str(p)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    n, s_, ws = r.stdout.split()
    en, es, ews = expected_sums(v, f)
    assert int(n) == en and len(f) > 100
    assert float(s_) == pytest.approx(es, rel=1e-12) and float(ws) == pytest.approx(ews, rel=1e-12)


def test_ply_loader_reports_truncated_and_inconsistent_files(host_bins, tmp_path):
    d = np.load(GOLDEN / "dragon_mesh.npz")
    v, f = d["vertices"][:300], d["faces"]
    f = [list(t) for t in f[(f < 300).all(1)][:200]]
    good = tmp_path / "good.ply"
    _write_general_ply(good, v, f, "binary_little_endian")
    raw = good.read_bytes()
    for cut in (len(raw) - 5, len(raw) // 2, raw.index(b"end_header") + 20):
        bad = tmp_path / "cut.ply"; bad.write_bytes(raw[:cut])
        r = subprocess.run([str(host_bins / "test_host_api"), "ply", str(bad)], capture_output=True, text=True)
        assert r.returncode == 1 and "truncated" in r.stdout, (cut, r.stdout)
    bad = tmp_path / "range.ply"
    _write_general_ply(bad, v, f + [[0, 1, 300]], "ascii", extras=False)                       # index == number of vertices
    r = subprocess.run([str(host_bins / "test_host_api"), "ply", str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and "out of range" in r.stdout
    for text in ("ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nproperty float y\nend_header\n0 0\n",      # no z
                 "ply\nformat ascii 1.0\nelement vertex 1\nproperty quaternion x\nend_header\n",                         # unknown type
                 "ply\nformat binary_middle_endian 1.0\nelement vertex 0\nproperty float x\nproperty float y\nproperty float z\nend_header\n",
                 "ply\nformat ascii 1.0\nelement vertex 0\nproperty float x\n"):                                          # no end_header
        bad = tmp_path / "hdr.ply"; bad.write_text(text)
        r = subprocess.run([str(host_bins / "test_host_api"), "ply", str(bad)], capture_output=True, text=True)
        assert r.returncode == 1 and "ERROR" in r.stdout, text


def test_ply_loader_skips_non_triangles_and_rejects_garbage(host_bins, tmp_path):
    p = tmp_path / "quad.ply"
    p.write_text("ply\nformat ascii 1.0\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\n"
                 "element face 2\nproperty list uchar int vertex_indices\nend_header\n"
                 "0 0 0\n1 0 0\n1 1 0\n0 1 0\n4 0 1 2 3\n3 0 1 2\n")
    r = subprocess.run([str(host_bins / "test_host_api"), "ply", str(p)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split()[0] == "1"
    bad = tmp_path / "bad.ply"; bad.write_text("not a ply\n")
    r = subprocess.run([str(host_bins / "test_host_api"), "ply", str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and "ERROR" in r.stdout
    r = subprocess.run([str(host_bins / "test_host_api"), "ply", str(tmp_path / "missing.ply")], capture_output=True, text=True)
    assert r.returncode == 1


@pytest.mark.gpu
def test_cli_and_render_loop_reproduce_the_reference_output(host_bins, tmp_path, oracle):
    """srt_xray (flags of main-pthreads-lbuffer.cxx) and srt::renderLoop write the text image the reference would:
    oracle repair of the reference-generated golden L-buffer, printed by Image::saveTextFile."""
    d = np.load(GOLDEN / "dragon_mesh.npz")
    write_ply(tmp_path / "dragon.ply", d["vertices"], d["faces"])
    (tmp_path / "out").mkdir()
    g = load_golden("lbuffer_dragon_128.npz")
    want = tmp_path / "want.txt"
    oracle.save_text(want, oracle.repair(g["L"], 128, 128), 128, 128)
    r = subprocess.run([str(host_bins / "srt_xray"), "-s", "128", "128", "-t", "6", "-f", "cli.txt", "-i", "./dragon.ply"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "out" / "cli.txt").read_text() == want.read_text()
    r = subprocess.run([str(host_bins / "test_host_api"), "gpu", str(tmp_path / "dragon.ply"), "128", "128",
                        str(tmp_path / "loop.txt"), "lbuffer"], capture_output=True, text=True)
    assert r.returncode == 0 and "22866 triangles" in r.stdout, r.stdout
    assert (tmp_path / "loop.txt").read_text() == want.read_text()
    # srt::renderLoop(Image&, meshes, info) is main.cxx's renderLoop: its text output must be the reference's own golden
    # file out/dragon-128x128-serial.txt, character for character (fixture: the file's values, printed with %g like
    # Image::saveTextFile prints them)
    r = subprocess.run([str(host_bins / "test_host_api"), "gpu", str(tmp_path / "dragon.ply"), "128", "128",
                        str(tmp_path / "serial.txt")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    gold = np.load(GOLDEN / "dragon_128_serial.npy")
    gold_text = "\n".join("\t".join("%.6g" % v for v in row) for row in gold)
    assert (tmp_path / "serial.txt").read_text() == gold_text
    # CLI: -b is accepted like the reference's, --csv appends a row in the batch scripts' format
    r = subprocess.run([str(host_bins / "srt_xray"), "-s", "64", "64", "-b", "10", "20", "30", "-f", "b.txt", "-i", "./dragon.ply",
                        "--csv", "timing.csv"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = (tmp_path / "timing.csv").read_text().strip().split("\n")
    assert rows[0] == "CPU,Parallelisation,Number of threads,Number of nodes,Compiler,Image size,Runtime in sec"
    assert len(rows) == 2 and rows[1].split(",")[1:4] == ["CUDA", "1", "1"] and rows[1].split(",")[5] == "64x64"
    assert subprocess.run([str(host_bins / "srt_xray"), "-b", "300", "0", "0"], cwd=tmp_path, capture_output=True).returncode != 0
    # the classic flag errors keep their exit status
    assert subprocess.run([str(host_bins / "srt_xray"), "--bogus"], capture_output=True).returncode != 0


@pytest.mark.gpu
def test_shaded_render_loop_cpp(host_bins, tmp_path):
    """srt::renderLoop(ImageRGB&, meshes, detector, origin, up, right, light): the main-omp.cxx flow in C++."""
    from simpleraytracing_b200 import scene
    d = np.load(GOLDEN / "dragon_mesh.npz")
    write_ply(tmp_path / "dragon.ply", d["vertices"], d["faces"])
    tex = scene.noise_texture()
    with open(tmp_path / "tex.ppm", "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (tex.shape[1], tex.shape[0])); f.write(tex.tobytes())
    r = subprocess.run([str(host_bins / "test_host_api"), "shaded", str(tmp_path / "dragon.ply"), "128", "128",
                        str(tmp_path / "tex.ppm"), str(tmp_path / "out.ppm")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = (tmp_path / "out.ppm").read_bytes()
    got = np.frombuffer(raw[raw.index(b"255\n") + 4:], np.uint8)
    want = np.load(GOLDEN / "shaded_dragon_logo_128.npz")["rgb_noise"]
    diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert diff.max() <= 1 and (diff == 0).mean() >= 0.999


@pytest.mark.gpu
def test_jpeg_writer_matches_the_8bit_lut_image(tmp_path):
    """srt_write_jpeg (nvJPEG on the GPU; replaces Image::saveJPEGFile, src/Image.cxx:85-146): the decoded file must be
    the 8-bit LUT image up to JPEG quantisation at quality 100 / 4:4:4 (mean error < 1 level, no pixel off by more than 8)."""
    from PIL import Image as PILImage
    from simpleraytracing_b200 import Context, Library, scene
    lib = Library(); ctx = Context(0, lib)
    tri = scene.dragon()
    lo, hi = lib.scene_bbox([(tri, 0.1037)])
    ctx.set_meshes([(tri, 0.1037)]); ctx.set_camera(lib.camera_from_bbox(lo, hi, 320, 200))
    img = ctx.render_xray(lib.default_params()).reshape(200, 320)
    lut = np.round(255.0 * np.clip(np.nan_to_num(img, nan=0.0), 0, 80) / 80.0).astype(np.uint8)      # applyLUT(0, 80)
    rgb = np.repeat(lut[:, :, None], 3, axis=2)
    ctx.write_jpeg(tmp_path / "xray.jpg", rgb, quality=100)
    back = np.asarray(PILImage.open(tmp_path / "xray.jpg").convert("RGB"))
    assert back.shape == rgb.shape
    d = np.abs(back.astype(np.int32) - rgb.astype(np.int32))
    assert d.mean() < 1.0 and d.max() <= 8, (d.mean(), d.max())
    # a colour image (the shaded variant's output format)
    col = scene.noise_texture(64, 96)
    ctx.write_jpeg(tmp_path / "c.jpg", col, quality=100)
    back = np.asarray(PILImage.open(tmp_path / "c.jpg").convert("RGB"))
    assert np.abs(back.astype(np.int32) - col.astype(np.int32)).mean() < 2.0
    ctx.close()
