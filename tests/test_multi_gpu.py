"""One process, N GPUs through the C ABI (srt_multi_*, include/srt.h): the role of the reference's MPI driver
(src/main-mpi.cxx:258-293 partition, :854-881 gather).  Every GPU renders its block-cyclic share and stores it into the
frame of every GPU over NVLink peer mappings; the image must equal the single-GPU render bit for bit.
With one visible GPU the same entry points run with N = 1; the peer paths need >= 2 GPUs (skipped otherwise)."""
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, bits
from simpleraytracing_b200 import KERNEL_CULLED, KERNEL_FILTERED, MultiContext, scene

pytestmark = pytest.mark.gpu
MU0 = float(scene.MU_MESH0)


def n_gpus():
    import torch
    return torch.cuda.device_count()


def single_image(gpu_ctx, srtlib, meshes, cam, kernel):
    gpu_ctx.set_meshes(meshes); gpu_ctx.set_camera(cam)
    return gpu_ctx.render_xray(srtlib.default_params(kernel=kernel))


@pytest.mark.parametrize("kernel", [KERNEL_FILTERED, KERNEL_CULLED])
def test_multi_context_on_one_gpu_equals_the_plain_context(gpu_ctx, srtlib, dragon, kernel):
    lo, hi = srtlib.scene_bbox([(dragon, MU0)])
    cam = srtlib.camera_from_bbox(lo, hi, 300, 150)
    want = single_image(gpu_ctx, srtlib, [(dragon, MU0)], cam, kernel)
    m = MultiContext([0], srtlib)
    m.set_meshes([(dragon, MU0)]); m.set_camera(cam)
    for _ in range(3):                                     # both frame sets
        got = m.render_xray(srtlib.default_params(kernel=kernel))
        assert (bits(got) == bits(want)).all()
    per, step = m.timing()
    assert per.shape == (1,) and per[0] > 0 and step >= per[0]
    m.close()


@pytest.mark.parametrize("kernel", [KERNEL_FILTERED, KERNEL_CULLED])
def test_all_visible_gpus_assemble_the_single_gpu_image(gpu_ctx, srtlib, dragon, kernel):
    n = n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs with peer access")
    import torch
    lo, hi = srtlib.scene_bbox([(dragon, MU0)])
    m = MultiContext(list(range(n)), srtlib)
    m.set_meshes([(dragon, MU0)])
    for W, H in ((2048, 2048), (300, 150), (1000, 810)):
        cam = srtlib.camera_from_bbox(lo, hi, W, H)
        want = single_image(gpu_ctx, srtlib, [(dragon, MU0)], cam, kernel)
        m.set_camera(cam)
        p = srtlib.default_params(kernel=kernel)
        for rep in range(4):                               # both frame sets, twice
            got = m.render_xray(p)
            assert (bits(got) == bits(want)).all(), (W, H, rep)
        # every device holds the complete repaired image and the complete pre-repair L-buffer
        m.render_xray_device(p); m.synchronize()
        ref = torch.from_numpy(want.view(np.int32))
        for i in range(n):
            dL, dimg = m.frame(i)
            img = _tensor_from_ptr(dimg, W * H, i)                     # torch view of the raw device pointer
            assert torch.equal(img.cpu().view(torch.int32), ref), (W, H, i)
    per, step = m.timing()
    assert len(per) == n and (per > 0).all()
    m.close()


def _tensor_from_ptr(ptr, n, device):
    """float32 torch view of n elements at a raw device pointer (test helper)."""
    import torch

    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
    with torch.cuda.device(device):
        return torch.as_tensor(h, device=f"cuda:{device}")


def test_multi_context_rejects_bad_device_lists(srtlib):
    from simpleraytracing_b200 import SrtError
    with pytest.raises(SrtError):
        MultiContext([], srtlib)
    with pytest.raises(SrtError):
        MultiContext([0, 0], srtlib)
    with pytest.raises(SrtError):
        MultiContext([999], srtlib)


def test_cpp_multi_gpu_renderer_equals_single_gpu(tmp_path):
    """host/test_host_api gpu-multi: srt::MultiGpuXRayRenderer on all visible GPUs vs srt::renderLoop on GPU 0."""
    from test_cpp_host import write_ply
    from conftest import ROOT
    subprocess.run(["make", "-C", str(ROOT / "simpleraytracing_b200" / "host")], check=True, capture_output=True)
    d = np.load(GOLDEN / "dragon_mesh.npz")
    write_ply(tmp_path / "dragon.ply", d["vertices"], d["faces"])
    r = subprocess.run([str(ROOT / "simpleraytracing_b200" / "bin" / "test_host_api"), "gpu-multi", str(tmp_path / "dragon.ply"),
                        "512", "512", str(tmp_path / "multi.txt")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
    assert f"{max(1, min(8, n_gpus()))} GPUs, 0 differing pixels" in r.stdout
