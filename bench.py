#!/usr/bin/env python3
"""bench.py -- headline benchmark of the X-ray L-buffer hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c4|c5] [--impl ours|reference]
                  [--lanes L] [--frames-per-step B] [--l-exchange halo|all] [--assembly auto|p2p|allgather]

Workloads (BASELINE.json configs):
  c3 (default, the configuration the metric is quoted on): dragon (22 866 triangles), 2048 x 2048
  c2: dragon, 1024 x 1024       c4: logo quad only (2 triangles), 1512 x 1512 (latency / memory bound)
  c5: synthetic 998 000-triangle sphere, 8192 x 8192
A FRAME is: L-buffer render + Beer-Lambert attenuation (block-cyclic 16-row blocks over the ranks at N > 1: L halo rows
to the neighbouring ranks by peer stores, device barrier, sharded artefact repair whose rows go to every rank's image
over NVLink / NVSwitch multicast, device barrier) -> the repaired image on every rank.
A STEP is a batch of B frames of the workload, issued round-robin on L lanes (an srt context + a stream + buffers each)
so that the latency-bound tail of one frame overlaps the scan of the next; every frame does all of its work inside the
timed region.  The frame is a fixed amount of work sharded over the ranks: scaling = "strong".

Keys of the JSON line (see the task contract):
  value        whole-job ray-triangle tests/s (tests of B x K frames / time of K steps), scene + frame buffers resident
               in HBM, brute-force kernel (SRT_KERNEL_FILTERED: every box of pixels meets every triangle, every
               (ray, triangle) pair is decided by a proven-conservative reject or the reference's exact sequence)
  ms_per_frame = ms_per_step / frames_per_step;  single_frame_ms = one frame with no other frame in flight
  e2e          N = 1: one frame through the host-pointer plugin call a reference user makes --
               srt_set_meshes(host vertices) + srt_set_camera + srt_render_xray(ordinary host image buffer) --
               timed by the wall clock around the synchronous calls; `pinned` repeats it with srt_host_alloc
               buffers.  N > 1 (one process per GPU): upload + sharded render + assembly + repair + image to host.
  roofline     the kernels are bound by the SM's instruction issue rate (no tensor-core or HBM roof applies):
               achieved = warp instructions / s of the scan kernel (instructions per test from the committed ncu
               capture x tests of the launch / kernel time by CUDA events, the kernel alone), peak = 4 issue slots x
               SMs x the SM clock sampled under load; `frame` = the same for all kernels of a frame over the measured
               pipelined time per frame.  The FP32 figure SURVEY 8d asks for is kept as `fp32_equivalent`.
  cpu_baseline the reference's CPU implementation (oracle/_ref, else the oracle port) on a bounded sample of the same
               frame on the host cores (rank 0, N = 1 only)
  accelerated  the same frames through SRT_KERNEL_CULLED (hierarchical conservative culling; identical results; what
               SRT_KERNEL_AUTO selects for meshes of more than ~2e5 triangles)

--impl reference times that CPU implementation as the bench arm itself.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

F_ALG = 19.0            # algorithmic flop per test
COUNTERS = ROOT / "profiles" / "r02_kernel_counters.json"      # ncu-derived per-kernel counters (tools/ncu_counters.py)


VERBOSE = bool(os.environ.get("SRT_BENCH_VERBOSE"))
T0 = time.time()


def note(msg):
    """progress marker on stderr (SRT_BENCH_VERBOSE=1): rank 0 only"""
    if VERBOSE and os.environ.get("RANK", "0") == "0":
        print(f"[bench {time.time() - T0:7.2f}s] {msg}", file=sys.stderr, flush=True)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ----------------------------------------------------------------------------- workloads
def workload(args):
    """-> dict(name, meshes [(tri, mu)], cam_tri (mesh the camera is derived from), W, H, tests per frame)."""
    from simpleraytracing_b200 import scene
    mu = float(scene.MU_MESH0)
    cfg = args.config
    if args.size:                                         # legacy: dragon at an arbitrary size
        tri = scene.dragon(); W = H = args.size
        name = f"X-ray L-buffer {W}x{H} dragon"
        return dict(name=name, meshes=[(tri, mu)], cam_tri=tri, W=W, H=H)
    if cfg in ("c2", "c3"):
        tri = scene.dragon(); W = H = 1024 if cfg == "c2" else 2048
        name = f"X-ray L-buffer {W}x{H} dragon (BASELINE configs[{1 if cfg == 'c2' else 2}])"
        return dict(name=name, meshes=[(tri, mu)], cam_tri=tri, W=W, H=H)
    if cfg == "c4":
        d = scene.dragon()
        lo = d.reshape(-1, 3).min(0); hi = d.reshape(-1, 3).max(0)
        quad = scene.logo_quad(lo, hi)
        return dict(name="logo-only attenuation 1512x1512, 2 triangles, camera of the dragon scene (BASELINE configs[3])",
                    meshes=[(quad, mu)], cam_tri=d, W=1512, H=1512)
    if cfg == "c5":
        tri = scene.uv_sphere()
        return dict(name="X-ray L-buffer 8192x8192 synthetic 998000-triangle sphere (BASELINE configs[4])",
                    meshes=[(tri, mu)], cam_tri=tri, W=8192, H=8192)
    raise SystemExit(f"unknown config {cfg}")


# ----------------------------------------------------------------------------- CPU arms
def cpu_reference_engine():
    """(kind, engine): the reference's own classes when oracle/_ref is built, else the oracle port."""
    from oracle.pyoracle import Oracle, Reference
    if Reference.available() or Path("/root/reference/src").is_dir():
        try:
            return "reference", Reference()
        except Exception:
            pass
    return "port", Oracle()


def cpu_sample_rate(engine, kind, wl, seconds, threads, n_pix=None):
    """Time the CPU L-buffer path on a contiguous pixel range through the frame centre.
    Brute-force cost is uniform per pixel, so tests/s of the sample is tests/s of the frame.
    n_pix = None: grow the sample until it takes about `seconds`; else time exactly that many pixels once."""
    tri = wl["meshes"][0][0]
    W, H = wl["W"], wl["H"]
    ntri = len(tri)
    cam = engine.camera(wl["cam_tri"], [len(wl["cam_tri"])]) if kind == "reference" else engine.camera(wl["cam_tri"])
    start = (H // 2) * W
    calibrate = n_pix is None
    if calibrate:
        n_pix = max(threads * 16, 256)
    while True:
        n_pix = int(min(W * H - start, n_pix))
        t0 = time.perf_counter()
        engine.render_lbuffer(tri, [ntri], cam, W, H, pixel_range=(start, start + n_pix - 1), threads=threads)
        dt = max(time.perf_counter() - t0, 1e-4)
        if not calibrate or dt >= 0.6 * seconds or n_pix >= W * H - start:
            return n_pix * ntri / dt, n_pix, dt
        n_pix = int(n_pix * min(8.0, max(1.5, 1.1 * seconds / dt)))


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    wl = workload(args)
    tri = wl["meshes"][0][0]
    W, H = wl["W"], wl["H"]
    kind, eng = cpu_reference_engine()
    threads = os.cpu_count() or 1
    # a step = one bounded sample of the frame: the sample size is calibrated once (about per_step seconds of work on all
    # host threads), every warm-up and timed step then renders exactly that pixel range
    per_step = max(0.5, min(8.0, 150.0 / max(1, args.steps + args.warmup)))
    _, n_pix, _ = cpu_sample_rate(eng, kind, wl, per_step, threads)
    for _ in range(args.warmup):
        cpu_sample_rate(eng, kind, wl, per_step, threads, n_pix)
    pix, secs = 0, 0.0
    for _ in range(args.steps):
        r, n, dt = cpu_sample_rate(eng, kind, wl, per_step, threads, n_pix)
        pix += n; secs += dt
    value = pix * len(tri) / secs
    frame_ms = W * H * len(tri) / value * 1e3
    sample = f"{n_pix} contiguous pixels of the {W}x{H} frame per step, all {len(tri)} triangles"
    line = {
        "impl": "reference", "metric": "ray-triangle tests/s (X-ray L-buffer)", "value": value,
        "unit": "tests/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs * 1e3 / max(1, args.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "dragon mesh fixture (tests/golden/dragon_mesh.npz) / synthetic; CPU, no GPU used",
        "ms_per_frame": frame_ms,
        "config": {"workload": wl["name"], "triangles": len(tri),
                   "step": f"one bounded sample of the frame ({n_pix} of {W * H} pixels x all triangles) on {threads} host threads; "
                           "ms_per_step is the measured time of a sample, ms_per_frame the full frame extrapolated from it "
                           "(brute-force cost is uniform per pixel)"},
        "cpu_baseline": {"value": value, "unit": "tests/s", "cores": threads, "cpu_model": cpu_model(), "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- our arm
def load_counters(kernel_name, shape, W, world):
    """Per-test instruction count and per-launch DRAM traffic of the dominant kernel from the committed ncu capture
    (profiles/r02_kernel_counters.json, written by tools/ncu_counters.py from an `ncu --set full` report)."""
    try:
        doc = json.loads(COUNTERS.read_text())
    except Exception:
        return None
    best = None
    for e in doc.get("captures", []):
        if e.get("kernel") != kernel_name:
            continue
        if shape is not None and e.get("box") not in (None, list(shape)):
            continue
        best = e
        if e.get("width") == W and e.get("n_gpus", 1) == world:
            break
    return best


def run_ours(args):
    # one hardware work queue per stream (the default of 8 would make the pipeline's lanes share queues and serialise)
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    import torch
    import torch.distributed as dist
    from simpleraytracing_b200 import Context, Library, PinnedArray, KERNEL_AUTO, KERNEL_CULLED, KERNEL_FILTERED
    from simpleraytracing_b200.bands import InterleavedLayout, PeerLanes
    from simpleraytracing_b200.capi import Camera

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    lib = Library()
    # the main stream: events, L2 flush, NCCL, copies.  Frames run on the lanes' streams, forked from / joined to it.
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    wl = workload(args)
    meshes, W, H = wl["meshes"], wl["W"], wl["H"]
    ntri = sum(len(m[0]) for m in meshes[:1])                     # only mesh 0 is traversed (lbuffer.cxx:788)
    lo, hi = lib.scene_bbox([(wl["cam_tri"], 0.0)])
    cam = lib.camera_from_bbox(lo, hi, W, H)
    # a second camera for the multi-GPU check: detector moved and pixel spacing changed, same frame size
    cam_b = Camera.from_buffer_copy(cam)
    cam_b.detector[1] = cam.detector[1] + 7.5; cam_b.detector[2] = cam.detector[2] - 4.25
    cam_b.pixel_spacing[0] = cam.pixel_spacing[0] * 0.9; cam_b.pixel_spacing[1] = cam.pixel_spacing[1] * 0.9
    note("scene built")
    params = lib.default_params(kernel=args.kernel, pixels_per_thread=args.box)
    # the accelerated path reported next to the headline: the culled kernel (bins + CTA tiles + pixel runs; SRT_KERNEL_AUTO
    # picks it for large meshes), or what AUTO picks for the logo-only scene of C4 (the one-launch exact kernel)
    params_acc = lib.default_params(kernel=KERNEL_AUTO if args.config == "c4" else KERNEL_CULLED)

    # ---- the pipeline: frames go round-robin to `n_lanes` lanes = one srt context + one stream + frame buffers each.
    # Nothing is shared between the frames in flight, so the latency-bound tail of one frame (candidate stages,
    # hit-list resolve, device barrier, repair) overlaps the scan of the next.
    big = args.config == "c5"
    # (a share of a frame is short: more frames must be in flight to cover the fixed latencies of its ~8 launches)
    n_lanes = max(1, args.lanes if args.lanes else (1 if big else (4 if world <= 2 else 8)))
    B = max(1, args.frames_per_step if args.frames_per_step else (1 if big else 4 * n_lanes))

    # N > 1: block-cyclic 16-row blocks (contiguous bands are unbalanced: rows through the object cost more).
    # Preferred assembly: the render kernels store their rows straight into every rank's frame over NVLink (multicast
    # mapping when the switch offers one) + srt_peer_barrier_device; fallback: all_gather of slots + de-interleave.
    il = InterleavedLayout(H, W, world, rank)
    my_rows = len(il.rows_of(rank)) if world > 1 else H
    peers, assembly = None, "single GPU"
    if world > 1:
        assembly = "all_gather of slots + de-interleave kernel"
        if args.assembly != "allgather":
            try:
                peers = PeerLanes(H, W, n_lanes, use_multicast=(args.assembly != "p2p"))
                assembly = ("L rows: " + ("own frame + the repair stencil's halo rows to the two neighbouring ranks (peer stores)" if args.l_exchange == "halo"
                                          else "every row to every rank") +
                            "; repaired image rows (sharded repair): " + ("multimem.st to the NVSwitch multicast mapping of all images" if peers.multicast_ptr
                                                                          else "peer stores to every rank's image") +
                            "; two flag-barrier kernels per frame")
            except Exception as exc:                      # no peer access on this box
                peers = None
                assembly += f" (symmetric memory unavailable: {type(exc).__name__})"
            # every rank must take the same path (the two assemblies synchronise differently): fall back everywhere
            # if any rank could not map its peers
            ok = torch.tensor([1 if peers is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0 and peers is not None:
                peers = None
                assembly = "all_gather of slots + de-interleave kernel (symmetric memory unavailable on another rank)"
        if peers is None:
            n_lanes = 1                                   # NCCL collectives stay on one stream

    class Lane:
        pass

    lanes = []
    for l in range(n_lanes):
        ln = Lane()
        ln.index = l
        ln.ctx = Context(local, lib)
        ln.stream = torch.cuda.Stream(device=dev) if n_lanes > 1 else stream
        ln.ctx.set_stream(ln.stream.cuda_stream)
        ln.ctx.set_meshes(meshes); ln.ctx.set_camera(cam)
        ln.ctx.set_profiling(False)
        if args.scan_ctas:
            ln.ctx.set_scan_residency(args.scan_ctas)
        ln.image = torch.empty(H * W, dtype=torch.float32, device=dev)
        ln.done = torch.cuda.Event()
        ln.epoch = 0
        if world == 1:
            # the same fused call as a rank of a sharded frame (one part: no barriers); its launches become a CUDA graph
            ln.frame = torch.empty(H * W, dtype=torch.float32, device=dev)
            ln.calls = {name: Context.frame_call_args(p_, [ln.frame.data_ptr()], [], [ln.image.data_ptr()])
                        for name, p_ in (("headline", params), ("acc", params_acc))}
        elif peers is not None:
            ln.calls = {}
            for name, p_ in (("headline", params), ("acc", params_acc)):
                ln.calls[name] = []
                for s_ in (0, 1):
                    img_ptrs, mc = peers.image_ptrs(l, s_)
                    if args.l_exchange == "halo":             # L rows: own frame + halo rows to the two neighbouring ranks
                        ln.calls[name].append(Context.frame_call_args(p_, peers.unicast_ptrs(l, s_), peers.flag_ptrs(l), img_ptrs,
                                                                      multicast=mc, halo=True))
                    else:                                     # every L row to every rank
                        ln.calls[name].append(Context.frame_call_args(p_, peers.frame_ptrs(l, s_)[0], peers.flag_ptrs(l), img_ptrs, multicast=mc))
            ln.own = [peers.own_frame_ptr(l, s_) for s_ in (0, 1)]
            ln.images = [peers.image(l, s_) for s_ in (0, 1)]
        else:
            ln.slots = il.new_slots(torch.float32, dev)
            ln.frame = torch.empty(H * W, dtype=torch.float32, device=dev)
        lanes.append(ln)
    ctx = lanes[0].ctx
    host_image_t = torch.empty(H * W, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def lane_frame(ln, which="headline"):
        """enqueue ONE frame (render of this rank's blocks + assembly + repair) on the lane's stream"""
        p = params if which == "headline" else params_acc
        if world == 1:
            ln.epoch += 1
            ln.ctx.render_xray_blocks_device(ln.calls[which], 1, 0, ln.epoch, ln.frame.data_ptr())
        elif peers is not None:
            ln.epoch += 1
            s_ = ln.epoch & 1
            ln.ctx.render_xray_blocks_device(ln.calls[which][s_], world, rank, ln.epoch, ln.own[s_])
            ln.image = ln.images[s_]                      # complete on every rank when the frame's second barrier has passed
        else:
            ln.ctx.render_lbuffer_interleaved_device(p, world, rank, il.my_slot(ln.slots).data_ptr())
            il.all_gather_(ln.slots)
            ln.ctx.deinterleave_device(world, ln.slots.data_ptr(), ln.frame.data_ptr())
            ln.ctx.repair_device(p, ln.frame.data_ptr(), ln.image.data_ptr())

    fork = torch.cuda.Event()

    def step_batch(which="headline", frames=None):
        """one step = `frames` frames, round-robin over the lanes; forked from and joined to the main stream"""
        frames = B if frames is None else frames
        if n_lanes > 1:
            fork.record(stream)
            for ln in lanes[:min(frames, n_lanes)]:
                ln.stream.wait_event(fork)
        for f in range(frames):
            lane_frame(lanes[f % n_lanes], which)
        if n_lanes > 1:
            for ln in lanes[:min(frames, n_lanes)]:
                ln.done.record(ln.stream)
                stream.wait_event(ln.done)

    def step_e2e_sharded(which="headline"):
        # one process per GPU: host meshes + camera in, sharded render, assembly, repair, image to host memory
        ln = lanes[0]
        ln.ctx.set_meshes(meshes)
        ln.ctx.set_camera(cam)
        if n_lanes > 1:
            fork.record(stream); ln.stream.wait_event(fork)
        lane_frame(ln, which)
        if n_lanes > 1:
            ln.done.record(ln.stream); stream.wait_event(ln.done)
        if rank == 0:                           # the result is needed on the host once
            host_image_t.copy_(ln.image, non_blocking=True)
        stream.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    per_rank_ms = []          # one list per timed() call: mean ms per step on every rank

    def timed(fn, steps, warmup, sync_each):
        for i in range(warmup):
            fn()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        wall = 0.0
        for a, b in ev:
            flush.fill_(1)                      # L2 flush between timed iterations (outside the event pair)
            if sync_each:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            a.record(stream)
            fn()
            b.record(stream)
            if sync_each:
                torch.cuda.synchronize()
                wall += time.perf_counter() - t0
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        if sync_each:
            ms = max(ms, wall * 1e3)            # host work (upload, pre-pack launch) is part of the end-to-end time
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            every = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(every, t)
            per_rank_ms.append([float(v.item()) / steps for v in every])
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def wall_timed(fn, steps, warmup):
        """Synchronous host-pointer calls: wall clock around each call (L2 flushed before, outside the clock)."""
        for _ in range(warmup):
            fn()
        total = 0.0
        for _ in range(steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            total += time.perf_counter() - t0
        return total * 1e3

    peak_tf = ctx.ffma_peak_tflops(3)                     # FP32 FFMA peak of this GPU (secondary figure)
    steps = args.steps
    # every lane renders once before anything is timed: the first render of a launch geometry verifies the list sizes
    for which in ("headline", "acc"):
        step_batch(which, frames=2 * n_lanes)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    note("timing the device-resident steps")
    total_ms = timed(step_batch, steps, args.warmup, sync_each=False)
    note(f"{total_ms / steps:.4f} ms per step of {B} frames")
    # one frame alone (no pipelining): the latency of a frame, and what a step was before frames were pipelined
    single_ms = timed(lambda: step_batch("headline", frames=1), max(3, min(steps, 10)), 2, sync_each=False)
    single_steps = max(3, min(steps, 10))
    # the dominant kernel alone: CUDA events around it on its stream (srt_timing.main_kernel_ms)
    barrier()
    ctx.set_profiling(True)
    kern_ms, pass_ms = [], []
    for _ in range(max(3, min(steps, 10))):
        flush.fill_(1)
        barrier()                                  # the flush has finished (it runs on the main stream, the frame on the lane's)
        lane_frame(lanes[0])
        tm = ctx.timing()
        kern_ms.append(tm.main_kernel_ms); pass_ms.append(tm.render_ms)
        render_launches = tm.launches              # dominant kernel + candidate stages + hit-list resolve (resp. bin pass)
        scan_cols = tm.scan_cols
    ctx.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    kernel_used = ctx.timing().kernel_used
    e2e_steps = max(1, min(steps, 10))

    # ---- end to end
    note("end to end")
    e2e = {}
    if world == 1:
        ectx = Context(local, lib)                                     # the plugin call as a reference user makes it: own stream
        host_image = np.empty(H * W, np.float32)                      # an ordinary (pageable) host buffer
        host_meshes = [(np.array(m[0], np.float32, copy=True), m[1]) for m in meshes]

        def plugin_call(img, ms_, p_=params):
            ectx.set_meshes(ms_)
            ectx.set_camera(cam)
            ectx.render_xray(p_, out=img)                             # synchronous: returns with the image in host memory

        e2e_ms = wall_timed(lambda: plugin_call(host_image, host_meshes), e2e_steps, 2)
        # the frame reaches a pageable buffer through a pinned staging area + copy threads (srt_set_host_copy_threads,
        # default = from the core count); the same call with the driver's own pageable copy (0) and other thread counts
        copy_sweep = {}
        for threads in (0, 1, 2, 4, 7):
            ectx.set_host_copy_threads(threads)
            copy_sweep[str(threads)] = wall_timed(lambda: plugin_call(host_image, host_meshes), e2e_steps, 2) / e2e_steps
        ectx.set_host_copy_threads(-1)
        e2e["pageable_copy"] = {"ms_per_step_by_helper_threads": copy_sweep, "default_threads": "a quarter of the host cores, at most 4",
                                "note": "0 = cudaMemcpyAsync into the pageable buffer (the driver stages it on one thread)"}
        pin_img = PinnedArray(lib, (H * W,), np.float32)
        pin_tri = [PinnedArray(lib, m[0].shape, np.float32) for m in meshes]
        for pa, m in zip(pin_tri, meshes):
            pa.array[...] = m[0]
        pinned_meshes = [(pa.array, m[1]) for pa, m in zip(pin_tri, meshes)]
        e2e_pin_ms = wall_timed(lambda: plugin_call(pin_img.array, pinned_meshes), e2e_steps, 2)
        e2e_same = bool(np.array_equal(pin_img.array.view(np.uint32), host_image.view(np.uint32)))
        acc_e2e_ms = wall_timed(lambda: plugin_call(pin_img.array, pinned_meshes, params_acc), e2e_steps, 2)
        e2e_path = "srt_set_meshes + srt_set_camera + srt_render_xray, ordinary host buffers, wall clock, one frame per call"
        e2e.update({"pinned": {"ms_per_step": e2e_pin_ms / e2e_steps, "value": float(W) * H * ntri * e2e_steps / (e2e_pin_ms * 1e-3),
                               "note": "same calls with srt_host_alloc (page-locked) vertex and image buffers",
                               "image_equals_pageable_run": e2e_same}})
        ectx.close()
    else:
        e2e_ms = timed(step_e2e_sharded, e2e_steps, max(1, min(args.warmup, 3)), sync_each=True)
        acc_e2e_ms = timed(lambda: step_e2e_sharded("acc"), e2e_steps, 2, sync_each=True)
        e2e_path = "per rank: srt_set_meshes + srt_set_camera + sharded render + assembly + repair; image to pinned host memory on rank 0; one frame per step"

    # the accelerated path (not the headline): same frames, same step, SRT_KERNEL_AUTO
    note("accelerated path")
    for ln in lanes:
        ln.ctx.set_meshes(meshes); ln.ctx.set_camera(cam)
    step_batch("acc", frames=2 * n_lanes); step_batch("headline", frames=2 * n_lanes)
    acc_ms = timed(lambda: step_batch("acc"), steps, 3, sync_each=False)
    acc_single_ms = timed(lambda: step_batch("acc", frames=1), single_steps, 2, sync_each=False)
    ctx.set_profiling(True); lane_frame(lanes[0], "acc"); acc_kernel_used = ctx.timing().kernel_used; ctx.set_profiling(False)

    # the shaded variant of the same frame size (main-omp.cxx: dragon + textured quad, closest hit, Phong, shadow
    # rays, texture), through the host-pointer call; N = 1, dragon configs only.  Not part of the headline.
    shaded = None
    if world == 1 and args.config in ("c2", "c3") and not args.no_shaded:
        note("shaded variant")
        from simpleraytracing_b200 import scene as scene_
        sctx = Context(local, lib)
        smeshes = scene_.shaded_scene(texture=scene_.noise_texture())
        scam, slight = scene_.shaded_camera(lib, smeshes, W, H)
        sctx.set_meshes(smeshes); sctx.set_camera(scam)
        p_sh = lib.default_params()
        sctx.render_shaded(p_sh, slight)
        first = sctx.timing()
        rows_ = []
        for _ in range(5):
            flush.fill_(1); torch.cuda.synchronize()
            t0_ = time.perf_counter()
            sctx.render_shaded(p_sh, slight)
            wall = (time.perf_counter() - t0_) * 1e3
            t_ = sctx.timing()
            rows_.append((wall, t_.render_ms, t_.repair_ms))
        brute = None
        if W * H <= 1024 * 1024 or args.shaded_brute_force:
            sctx.force_shadow_grid(-1)
            sctx.render_shaded(p_sh, slight)
            brute = sctx.timing().repair_ms
        med = np.median(np.array(rows_), axis=0)
        shaded = {"workload": f"shaded {W}x{H}: dragon + textured quad ({sum(len(m['tri']) for m in smeshes)} triangles), light of main-omp.cxx",
                  "e2e_ms_per_frame": float(med[0]), "closest_hit_pass_ms": float(med[1]), "shade_pass_ms": float(med[2]),
                  "shadow_grid": {"cells_per_side": int(first.shadow_grid_cells), "list_entries": int(first.shadow_grid_entries),
                                  "build_ms_first_frame": float(first.shadow_grid_ms)},
                  "shade_pass_ms_without_grid": brute,
                  "d2h_bytes_per_frame": 3 * W * H}
        sctx.close()

    # ---- correctness of the assembled frames (outside every timed region).  The image a rank holds after a sharded
    # frame must equal, bit for bit, the image of one unsharded render on this GPU.  Each round of the check gives the
    # lanes alternating cameras, poisons every frame buffer of every rank with NaN, then runs a PIPELINED batch of
    # 2 x lanes frames (both frame sets of every lane, all lanes in flight) -- a lost or late peer store, or a barrier
    # that lets the repair pass start early, cannot hide behind bytes left by an earlier frame.
    verified, checks = True, 0
    if world > 1:
        vctx = Context(local, lib)
        vctx.set_stream(stream.cuda_stream)
        vctx.set_meshes(meshes)
        ref_L = torch.empty(H * W, dtype=torch.float32, device=dev)
        ref_img = {}
        rounds = max(1, args.verify_iters // (2 * n_lanes))
        for it in range(rounds):
            which = "acc" if (it % 3 == 2) else "headline"
            p_ = params_acc if which == "acc" else params
            cams = [cam_b if ((ln.index + it) & 1) else cam for ln in lanes]
            for ln, c_ in zip(lanes, cams):
                ln.ctx.set_camera(c_)
                ln.image.fill_(float("nan"))
            for key, c_ in (("a", cam), ("b", cam_b)):
                vctx.set_camera(c_)
                vctx.render_lbuffer_device(p_, (0, H), ref_L.data_ptr())
                ref_img[key] = torch.empty(H * W, dtype=torch.float32, device=dev)
                vctx.repair_device(p_, ref_L.data_ptr(), ref_img[key].data_ptr())
            # three unchecked batches first: from its third occurrence a call signature is replayed from a captured CUDA graph
            # (srt_render_xray_blocks_device), and the checked batch must be one of those
            for _ in range(3):
                step_batch(which, frames=2 * n_lanes)
            barrier()
            if peers is not None:
                peers.poison()
            else:
                for ln in lanes:
                    ln.slots.fill_(float("nan")); ln.frame.fill_(float("nan"))
            barrier()                                           # every rank has poisoned its buffers
            step_batch(which, frames=2 * n_lanes)
            torch.cuda.synchronize()
            for ln, c_ in zip(lanes, cams):
                same = torch.equal(ln.image.view(torch.int32), ref_img["b" if c_ is cam_b else "a"].view(torch.int32))
                verified = verified and bool(same)
                checks += 2
            barrier()
        for ln in lanes:
            ln.ctx.set_camera(cam)
        vctx.close()
        flag = torch.tensor([1 if verified else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        verified = bool(flag.item())
        if not verified:
            raise SystemExit(f"bench.py: rank {rank}: an assembled frame differs from the unsharded render")
        step_batch("headline", frames=2 * n_lanes)              # back to the bench camera (list sizes re-verified)
        barrier()

    note("exact-test counts")
    # actual work behind the "tests": pairs that reach the exact reference sequence (this rank's share), both kernels
    exact_counts = {}
    try:
        ctx.count_exact(True)
        for name, which in (("headline", "headline"), ("accelerated", "acc")):
            if world > 1:
                barrier()
            lane_frame(lanes[0], which)
            exact_counts[name] = int(ctx.timing().exact_tests)
    except Exception:
        exact_counts = {}
    finally:
        ctx.count_exact(False)

    tests_frame = float(W) * H * ntri
    value = tests_frame * B * steps / (total_ms * 1e-3)
    e2e_value = tests_frame * e2e_steps / (e2e_ms * 1e-3)
    kt = float(np.mean(kern_ms)) * 1e-3
    per_rank_kernel = None
    if world > 1:
        kt_t = torch.tensor([float(np.mean(pass_ms))], dtype=torch.float64, device=dev)
        every = [torch.zeros_like(kt_t) for _ in range(world)]
        dist.all_gather(every, kt_t)
        per_rank_kernel = [round(float(v.item()), 4) for v in every]
    band_tests = float(my_rows) * W * ntri

    if rank == 0:
        cpu = None
        note("cpu baseline")
        if world == 1 and not args.no_cpu:
            kind, eng = cpu_reference_engine()
            threads = os.cpu_count() or 1
            r, n_pix, dt = cpu_sample_rate(eng, kind, wl, args.cpu_seconds, threads)
            cpu = {"value": r, "unit": "tests/s", "cores": threads, "cpu_model": cpu_model(), "kind": kind,
                   "sample": f"{n_pix} contiguous pixels of the {W}x{H} frame x {ntri} triangles in {dt:.1f} s"}
        # ---- roofline of the dominant kernel: instruction-issue bound
        kname = {KERNEL_FILTERED: "lbuffer_filtered_kernel", KERNEL_CULLED: "lbuffer_culled_kernel"}.get(kernel_used, "lbuffer_exact_kernel")
        cnt = load_counters(kname, None, W, world)
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        issue_peak = 4.0 * sms * sm_mhz * 1e6 / 1e9                      # G warp-instructions / s
        roof = {"bound": "issue", "unit": "Gwarp-inst/s", "peak": issue_peak,
                "peak_source": f"4 issue slots x {sms} SMs x {sm_mhz:.0f} MHz (SM clock sampled under load)",
                "kernel": kname, "kernel_ms": kt * 1e3, "pass_ms": float(np.mean(pass_ms)), "scan_cols": int(scan_cols),
                "kernel_timing": "the kernel alone on an otherwise idle GPU (one frame, CUDA events around the launch)",
                "tests_per_launch": band_tests, "exact_sequence_tests_per_launch": exact_counts.get("headline"),
                "achieved": None, "frac": None, "traffic": None}
        if cnt:
            ipt = float(cnt["warp_inst_per_test"])
            roof["achieved"] = ipt * band_tests / kt / 1e9
            roof["frac"] = roof["achieved"] / issue_peak
            roof["warp_inst_per_test"] = ipt
            roof["counters_source"] = f"profiles/{COUNTERS.name}: {cnt.get('source', '')}"
            if cnt.get("width") == W and cnt.get("n_gpus", 1) == world:
                roof["traffic"] = cnt.get("dram_bytes")              # dram__bytes_read.sum + dram__bytes_write.sum, this launch shape
            roof["algorithmic_bytes"] = 4.0 * my_rows * W             # the L-buffer rows this launch produces
        # the whole frame: the six kernels of a frame together (warp instructions from the same ncu capture) over the measured,
        # pipelined time per frame -- how much of the GPU's issue capacity the steady state uses
        try:
            caps = [c for c in json.loads(COUNTERS.read_text()).get("captures", [])
                    if c.get("width") == W and c.get("n_gpus", 1) == 1 and c.get("box") == (cnt or {}).get("box")]
            if world == 1 and kernel_used == KERNEL_FILTERED and caps:
                frame_inst = float(sum(c["warp_inst"] for c in caps))
                per_frame_s = total_ms / steps / B * 1e-3
                roof["frame"] = {"kernels": {c["kernel"]: {"us_cold_serialised": round(c["duration_us"], 1), "warp_inst": c["warp_inst"],
                                                           "ipc_per_sm": c["ipc_per_sm"]} for c in caps},
                                 "warp_inst_per_frame": frame_inst, "achieved": frame_inst / per_frame_s / 1e9,
                                 "frac": frame_inst / per_frame_s / 1e9 / issue_peak,
                                 "note": "all kernels of a frame / measured pipelined time per frame; unit and peak as above"}
                roof["kernel_share_of_frame"] = next((c["duration_us"] for c in caps if c["kernel"] == kname), 0.0) / sum(c["duration_us"] for c in caps)
        except Exception:
            pass
        # secondary: the reference's u-stage costed at 19 flop per test against the measured FFMA peak; the
        # kernel decides that stage with far fewer instructions, so this is an equivalence figure, not a fraction of peak
        roof["fp32_equivalent"] = {"flop_per_test": F_ALG, "equivalent_tflops": band_tests * F_ALG / kt / 1e12,
                                   "ffma_peak_tflops": peak_tf, "note": "equivalent rate / FFMA peak is NOT a utilisation"}
        h2d = int(sum(m[0].nbytes for m in meshes) + 4 * 16)
        per_frame_launches = render_launches + (1 if world == 1 else (3 if peers is not None else 2))   # + repair (+ 2 barrier kernels | de-interleave)
        line = {
            "metric": "ray-triangle tests/s (X-ray L-buffer)", "value": value, "unit": "tests/s",
            "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": total_ms / steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "dragon mesh fixture (reference dragon.ply, 22866 triangles) / synthetic sphere; camera from the scene bbox",
            "frames_per_step": B, "ms_per_frame": total_ms / steps / B,
            "single_frame_ms": single_ms / single_steps,
            "config": {"workload": wl["name"], "triangles": ntri, "kernel": kernel_used,
                       "kernel_note": "2 = SRT_KERNEL_FILTERED: the brute-force path, every thread visits every triangle",
                       "l2": "256 MiB flush write between steps",
                       "step": f"a batch of {B} frames of the workload, each = render (16-row blocks, block-cyclic over ranks) + assembly + repair; "
                               f"frames round-robin over {n_lanes} lanes (srt context + stream each) so that the tail of one frame overlaps the scan of the next; "
                               "single_frame_ms = one frame alone",
                       "lanes": n_lanes, "scan_ctas_per_sm_cap": args.scan_ctas,
                       "assembly": assembly,
                       "assembled_frame_equals_unsharded_render": verified,
                       "assembly_check": None if world == 1 else f"{checks} sharded frames in pipelined batches (graph-replayed), two cameras alternating over the lanes, frames poisoned with NaN before each checked batch"},
            "e2e": {"value": e2e_value, "unit": "tests/s", "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                    "path": e2e_path, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(H * W * 4), **e2e},
            "gpu_launches": per_frame_launches * B * steps,
            "roofline": roof,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "per_rank": None if world == 1 else {"render_pass_ms": per_rank_kernel,
                                                 "step_ms": [round(v, 4) for v in per_rank_ms[0]]},
            "shaded": shaded,
            "accelerated": {"kernel": acc_kernel_used,
                            "kernel_name": "4 = SRT_KERNEL_CULLED (bins + CTA tiles + pixel runs; what SRT_KERNEL_AUTO selects above ~2e5 triangles), 1 = exact kernel (AUTO for a handful of triangles); identical results",
                            "ms_per_step": acc_ms / steps, "ms_per_frame": acc_ms / steps / B, "single_frame_ms": acc_single_ms / single_steps,
                            "value": tests_frame * B * steps / (acc_ms * 1e-3),
                            "unit": "tests/s (pairs resolved, most of them by a conservative box test)",
                            "e2e_ms_per_step": acc_e2e_ms / e2e_steps,
                            "exact_sequence_tests_per_launch": exact_counts.get("accelerated")},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    for ln in lanes:
        ln.ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c3", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--size", type=int, default=0, help="dragon at SIZE x SIZE instead of a named config")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernel", type=int, default=2, help="headline kernel: 2 filtered (brute force, default), 1 exact, 4 culled, 0 auto")
    ap.add_argument("--box", type=lambda v: int(v, 0), default=0,
                    help="box shape of the filtered kernel: width | height << 8 | boxes per thread << 16 (0 = default)")
    ap.add_argument("--assembly", default="auto", choices=["auto", "p2p", "allgather"],
                    help="N > 1: auto = fused stores (multicast if available, else per-peer), p2p = per-peer stores, allgather = NCCL")
    ap.add_argument("--l-exchange", default="halo", choices=["halo", "all"],
                    help="N > 1: which L rows leave the rank: halo = only what the neighbours' repair stencils read, all = every row to every rank")
    ap.add_argument("--verify-iters", type=int, default=48, help="N > 1: sharded frames checked against unsharded renders")
    ap.add_argument("--lanes", type=int, default=0, help="frames in flight (srt context + stream each); 0 = default (4, 8 at N > 2; 1 for c5)")
    ap.add_argument("--frames-per-step", type=int, default=0, help="frames per timed step; 0 = default (4 x lanes; 1 for c5)")
    ap.add_argument("--scan-ctas", type=int, default=0, help="cap on the scan kernel's resident CTAs per SM (0 = none, else >= 4)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-shaded", action="store_true", help="skip the shaded-variant block (N = 1, dragon configs)")
    ap.add_argument("--shaded-brute-force", action="store_true",
                    help="also time the shade pass without the shadow grid at sizes above 1024^2 (seconds per frame)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.config == "c5" and args.steps > 5 and "--steps" not in " ".join(sys.argv):
        args.steps = 3                                           # seconds per frame in the brute-force kernel
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
