#!/usr/bin/env python3
"""Frame pipelining on ONE GPU: throughput of batches of frames rendered round-robin on L lanes (context + stream
each), for the whole frame and for ONE block-cyclic share of an N-GPU frame (no peers: the share's kernels + a full-frame
repair, i.e. the device work of a rank at N GPUs without the barrier).

  python tools/lanes.py [size] [frames per batch] [shape PXxPYxK ...]

Output: one line per (share, lanes, scan residency cap); copy what matters into profiles/."""
import os
import sys
import time
from pathlib import Path

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simpleraytracing_b200 import Context, Library, scene, KERNEL_FILTERED, KERNEL_AUTO  # noqa: E402
from simpleraytracing_b200.capi import box_shape  # noqa: E402


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    lib = Library()
    tri = scene.dragon()
    lo, hi = lib.scene_bbox([(tri, 0.1037)])
    cam = lib.camera_from_bbox(lo, hi, size, size)
    dev = torch.device("cuda", 0)
    main_s = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(main_s)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    max_lanes = 6
    ctxs, streams, frames, images = [], [], [], []
    for _ in range(max_lanes):
        c = Context(0, lib); s = torch.cuda.Stream(device=dev)
        c.set_stream(s.cuda_stream); c.set_meshes([(tri, 0.1037)]); c.set_camera(cam); c.set_profiling(False)
        ctxs.append(c); streams.append(s)
        frames.append(torch.zeros(size * size, dtype=torch.float32, device=dev)); images.append(torch.zeros(size * size, dtype=torch.float32, device=dev))
    fork = torch.cuda.Event(); done = [torch.cuda.Event() for _ in range(max_lanes)]

    def batch(p, n_parts, lanes, nframes):
        fork.record(main_s)
        for l in range(min(lanes, nframes)):
            streams[l].wait_event(fork)
        for f in range(nframes):
            l = f % lanes
            if n_parts == 1:
                ctxs[l].render_lbuffer_device(p, (0, size), frames[l].data_ptr())
            else:
                ctxs[l].render_lbuffer_blocks_device(p, n_parts, (3 * l + 1) % n_parts, [frames[l].data_ptr()])   # a fixed share per lane: list sizes are verified per launch geometry
            ctxs[l].repair_device(p, frames[l].data_ptr(), images[l].data_ptr())
        for l in range(min(lanes, nframes)):
            done[l].record(streams[l]); main_s.wait_event(done[l])

    def measure(p, n_parts, lanes, nframes, reps=6):
        for _ in range(2):
            batch(p, n_parts, lanes, max(nframes, 2 * lanes))
        torch.cuda.synchronize()
        ms, cpu = [], []
        for _ in range(reps):
            flush.fill_(1)
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            a.record(main_s); batch(p, n_parts, lanes, nframes); b.record(main_s)
            cpu.append((time.perf_counter() - t0) * 1e3)               # host time to enqueue the batch
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return min(ms) / nframes, sorted(ms)[len(ms) // 2] / nframes, min(cpu) / nframes

    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[3:]] or [(0, 0, 0)]
    runs = [(KERNEL_FILTERED, f"filtered {sh[0]}x{sh[1]}x{sh[2]}" if sh[0] else "filtered default", sh) for sh in shapes] + [(KERNEL_AUTO, "auto", (0, 0, 0))]
    for kern, kname, sh in runs:
        p = lib.default_params(kernel=kern, pixels_per_thread=box_shape(*sh) if sh[0] else 0)
        base = {}
        for n_parts in (1, 8):
            for cap in (0, 4):
                for c in ctxs:
                    c.set_scan_residency(cap)
                for lanes in (1, 2, 4, 6):
                    if kern != KERNEL_FILTERED and cap:
                        continue
                    one = measure(p, n_parts, lanes, 1)[1] if lanes == 1 else None
                    best, med, cpu = measure(p, n_parts, lanes, B)
                    if n_parts == 1 and lanes == 1 and cap == 0:
                        base[kname] = med
                    ideal = base.get(kname, med) / n_parts
                    print(f"{kname} {size}^2 share 1/{n_parts} lanes {lanes} scan cap {cap}: {med * 1e3:8.1f} us/frame (best {best * 1e3:8.1f})"
                          + (f" single frame {one * 1e3:8.1f}" if one is not None else "")
                          + f" | vs 1-lane full frame / {n_parts}: {ideal / med:.3f} | host enqueue {cpu * 1e3:.1f} us/frame", flush=True)
    for c in ctxs:
        c.close()


if __name__ == "__main__":
    main()
