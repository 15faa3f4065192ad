"""The sharded frame pipeline of bench.py (one process per GPU: srt_render_xray_blocks_device = render of the rank's
blocks stored into every rank's frame + srt_peer_barrier_device + repair), exercised on ONE GPU: every "rank" is a
context with its own stream in this process, the "peer-mapped" frames and flag arrays are plain buffers of the same
device.  The kernels, the barrier protocol (flags, epochs, two frame sets per lane) and the concurrency of several
lanes are exactly those of the multi-GPU run; only the transport differs.  Replaces the gather of
src/main-mpi.cxx:854-881.  The image must equal the unsharded render bit for bit."""
import numpy as np
import pytest

from conftest import bits
from simpleraytracing_b200 import Context, KERNEL_CULLED, KERNEL_FILTERED, scene

pytestmark = pytest.mark.gpu
MU0 = float(scene.MU_MESH0)
FLAG_WORDS = 64


class FakeRanks:
    """n_ranks x lanes contexts on cuda:0 with frames / images / flags laid out like bands.PeerLanes"""

    def __init__(self, srtlib, meshes, cam, n_ranks, lanes, params, halo=False):
        import torch
        self.torch = torch
        self.n, self.lanes, self.W, self.H = n_ranks, lanes, cam.width, cam.height
        px = self.px = self.W * self.H
        dev = torch.device("cuda", 0)
        # buf[rank][lane]: [L frame 0][L frame 1][image 0][image 1][flags]
        self.buf = torch.zeros(n_ranks, lanes, 4 * px + FLAG_WORDS, dtype=torch.float32, device=dev)
        self.ctx = [[Context(0, srtlib) for _ in range(lanes)] for _ in range(n_ranks)]
        self.streams = [[torch.cuda.Stream(device=dev) for _ in range(lanes)] for _ in range(n_ranks)]
        self.epoch = [0] * lanes
        self.params = params
        self.halo = halo                                 # SRT_FLAG_L_HALO_ONLY: L rows go to the neighbouring ranks only
        for r in range(n_ranks):
            for l in range(lanes):
                c = self.ctx[r][l]
                c.set_stream(self.streams[r][l].cuda_stream)
                c.set_meshes(meshes); c.set_camera(cam)
                c.set_profiling(False)

    def ptrs(self, r, l, which):
        """buffer `which` (0, 1: L frames; 2, 3: images) of lane l: rank r's own first, then the other ranks'"""
        px = self.px
        own = self.buf[r, l, which * px:(which + 1) * px].data_ptr()
        return [own] + [self.buf[q, l, which * px:(which + 1) * px].data_ptr() for q in range(self.n) if q != r]

    def image(self, r, l):
        s = self.epoch[l] & 1
        return self.buf[r, l, (2 + s) * self.px:(3 + s) * self.px]

    def frame(self, l, order):
        """one frame of lane `l` on every rank: srt_render_xray_blocks_device taken apart.  All ranks are driven by this
        one thread, and streams may share a hardware queue, so at both barriers every rank's signal is enqueued
        before the first wait (a real rank uses phases = 3)."""
        self.epoch[l] += 1
        e = self.epoch[l]
        s = e & 1
        flags = [self.buf[q, l, 4 * self.px:].data_ptr() for q in range(self.n)]
        for r in order:
            self.ctx[r][l].render_lbuffer_blocks_device(self.params, self.n, r, self.ptrs(r, l, s), halo=self.halo)
        for r in order:
            self.ctx[r][l].peer_barrier_device(self.n, r, flags, 2 * e - 1, phases=1)
        for r in order:
            self.ctx[r][l].peer_barrier_device(self.n, r, flags, 2 * e - 1, phases=2)
            self.ctx[r][l].repair_blocks_device(self.params, self.n, r, self.ptrs(r, l, s)[0], self.ptrs(r, l, 2 + s))
        for r in order:
            self.ctx[r][l].peer_barrier_device(self.n, r, flags, 2 * e, phases=1)
        for r in order:
            self.ctx[r][l].peer_barrier_device(self.n, r, flags, 2 * e, phases=2)

    def poison(self):
        self.buf[:, :, :4 * self.px].fill_(float("nan"))

    def close(self):
        self.torch.cuda.synchronize()
        for row in self.ctx:
            for c in row:
                c.close()


@pytest.mark.parametrize("kernel,n_ranks,lanes,halo", [(KERNEL_FILTERED, 2, 1, False), (KERNEL_FILTERED, 4, 3, True), (KERNEL_CULLED, 3, 2, True),
                                                       (KERNEL_FILTERED, 2, 2, True), (KERNEL_CULLED, 8, 1, True)])
def test_sharded_pipeline_equals_the_unsharded_render(gpu_ctx, srtlib, dragon, kernel, n_ranks, lanes, halo):
    import torch
    lo, hi = srtlib.scene_bbox([(dragon, MU0)])
    cam = srtlib.camera_from_bbox(lo, hi, 600, 410)
    p = srtlib.default_params(kernel=kernel)
    gpu_ctx.set_meshes([(dragon, MU0)]); gpu_ctx.set_camera(cam)
    want = torch.from_numpy(gpu_ctx.render_xray(p).view(np.int32)).cuda()
    fr = FakeRanks(srtlib, [(dragon, MU0)], cam, n_ranks, lanes, p, halo=halo)
    for rnd in range(3):
        fr.poison()
        torch.cuda.synchronize()
        # 2 frames per lane (both frame sets), all lanes and all ranks in flight at once
        order = list(range(n_ranks)) if rnd != 1 else list(reversed(range(n_ranks)))
        for f in range(2):
            for l in range(lanes):
                fr.frame(l, order)
        torch.cuda.synchronize()
        for r in range(n_ranks):
            for l in range(lanes):
                assert torch.equal(fr.image(r, l).view(torch.int32), want), (rnd, r, l)
    fr.close()


def test_profiling_off_and_scan_residency_do_not_change_the_image(gpu_ctx, srtlib, dragon):
    lo, hi = srtlib.scene_bbox([(dragon, MU0)])
    cam = srtlib.camera_from_bbox(lo, hi, 512, 512)
    p = srtlib.default_params(kernel=KERNEL_FILTERED)
    gpu_ctx.set_meshes([(dragon, MU0)]); gpu_ctx.set_camera(cam)
    want = gpu_ctx.render_xray(p)
    c = Context(0, srtlib)
    c.set_meshes([(dragon, MU0)]); c.set_camera(cam)
    c.set_profiling(False)
    for cap in (0, 4, 0):
        c.set_scan_residency(cap)
        for _ in range(3):                                   # the list counters alternate between their two pairs
            got = c.render_xray(p)
            assert (bits(got) == bits(want)).all(), cap
    t = c.timing()
    assert t.render_ms == 0.0 and t.main_kernel_ms == 0.0     # no events were recorded
    c.set_profiling(True)
    got = c.render_xray(p)
    assert (bits(got) == bits(want)).all() and c.timing().render_ms > 0.0
    from simpleraytracing_b200 import SrtError
    with pytest.raises(SrtError):
        c.set_scan_residency(2)
    c.close()


def test_fused_frame_call_with_one_rank_equals_render_xray(gpu_ctx, srtlib, dragon):
    import torch
    lo, hi = srtlib.scene_bbox([(dragon, MU0)])
    cam = srtlib.camera_from_bbox(lo, hi, 300, 150)
    p = srtlib.default_params(kernel=KERNEL_FILTERED)
    gpu_ctx.set_meshes([(dragon, MU0)]); gpu_ctx.set_camera(cam)
    want = torch.from_numpy(gpu_ctx.render_xray(p).view(np.int32)).cuda()
    frame = torch.zeros(300 * 150, dtype=torch.float32, device="cuda"); img = torch.zeros_like(frame)
    flags = torch.zeros(FLAG_WORDS, dtype=torch.int32, device="cuda")
    args = Context.frame_call_args(p, [frame.data_ptr()], [flags.data_ptr()], [img.data_ptr()])
    gpu_ctx.set_profiling(False)                             # steady-state calls are captured into a CUDA graph and replayed
    try:
        for e in range(1, 13):
            img.fill_(float("nan")); frame.fill_(float("nan"))
            gpu_ctx.synchronize(); torch.cuda.synchronize()
            gpu_ctx.render_xray_blocks_device(args, 1, 0, e, frame.data_ptr())
            gpu_ctx.synchronize()
            assert torch.equal(img.view(torch.int32), want), e
        # the sharded repair pass alone: three shares of the frame into one image == the full-frame repair (300 is not a
        # multiple of 4 x 256 threads, 150 rows leave a ragged last block)
        img.fill_(float("nan"))
        for part in range(3):
            gpu_ctx.repair_blocks_device(p, 3, part, frame.data_ptr(), [img.data_ptr()])
        gpu_ctx.synchronize()
        assert torch.equal(img.view(torch.int32), want)
        # a new camera: the graphs of the old one must not be replayed
        cam2 = srtlib.camera_from_bbox(lo, hi, 300, 150); cam2.pixel_spacing[0] *= 0.8; cam2.pixel_spacing[1] *= 0.8
        gpu_ctx.set_camera(cam2)
        want2 = torch.from_numpy(gpu_ctx.render_xray(p).view(np.int32)).cuda()
        assert not torch.equal(want2, want)
        for e in range(13, 21):
            img.fill_(float("nan"))
            torch.cuda.synchronize()
            gpu_ctx.render_xray_blocks_device(args, 1, 0, e, frame.data_ptr())
            gpu_ctx.synchronize()
            assert torch.equal(img.view(torch.int32), want2), e
        gpu_ctx.set_camera(cam)
    finally:
        gpu_ctx.set_profiling(True)
    from simpleraytracing_b200 import SrtError
    with pytest.raises(SrtError):
        gpu_ctx.peer_barrier_device(2, 0, [flags.data_ptr(), flags.data_ptr()], 1, phases=0)
    with pytest.raises(SrtError):
        gpu_ctx.repair_blocks_device(p, 3, 0, frame.data_ptr(), [frame.data_ptr()])           # image aliases the L frame
