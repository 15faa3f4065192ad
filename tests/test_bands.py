"""Multi-GPU assembly logic (row bands + in-place all_gather) on CPU: world_size 2 and 3 over gloo."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, H, W, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simpleraytracing_b200.bands import BandLayout
    lay = BandLayout(H, W, world, rank)
    full = torch.arange(H * W, dtype=torch.float32) * 0.5 + 1.0          # the frame every rank should end with
    frame = lay.new_frame(torch.float32, "cpu")
    rb, re = lay.rows
    frame[: H * W].view(H, W)[rb:re] = full.view(H, W)[rb:re]           # "render" my band in place
    out = lay.all_gather_(frame)
    ok = bool((out == full).all()) and out.numel() == H * W
    q.put((rank, ok, lay.rows))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,H,W", [(2, 10, 7), (2, 9, 5), (3, 10, 4), (2, 1, 3)])
def test_band_allgather_gloo(world, H, W):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, W, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res), res
    rows = sorted(r for _, _, r in res)
    assert rows[0][0] == 0 and rows[-1][1] == H and all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))


def _worker_il(rank, world, port, H, W, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simpleraytracing_b200.bands import InterleavedLayout
    lay = InterleavedLayout(H, W, world, rank)
    full = (torch.arange(H * W, dtype=torch.float32) * 0.25 - 3.0).view(H, W)
    slots = lay.new_slots(torch.float32, "cpu")
    rows = lay.rows_of(rank)
    if rows:
        lay.my_slot(slots).view(lay.slot_rows, W)[: len(rows)] = full[torch.tensor(rows)]    # "render" my blocks
    lay.all_gather_(slots)
    out = lay.deinterleave_torch(slots)
    q.put((rank, bool((out == full.view(-1)).all()), rows))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,H,W", [(2, 64, 5), (2, 40, 3), (3, 100, 4), (2, 7, 2)])
def test_interleaved_allgather_gloo(world, H, W):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_il, args=(r, world, port, H, W, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res), res
    assert sorted(sum((rows for _, _, rows in res), [])) == list(range(H))       # every row rendered exactly once


def test_interleaved_layout_matches_library():
    from simpleraytracing_b200.bands import InterleavedLayout
    from simpleraytracing_b200 import capi
    lib = capi.Library()
    for H, n in ((2048, 8), (2048, 3), (100, 4), (5, 8), (16, 1)):
        assert InterleavedLayout(H, 8, n, 0).slot_rows == lib.interleave_slot_rows(H, n)


def test_band_layout_matches_reference_partition():
    from simpleraytracing_b200.bands import BandLayout
    from simpleraytracing_b200 import capi
    lib = capi.Library()
    for H, n in ((2048, 8), (2048, 3), (10, 4), (5, 8)):
        for i in range(n):
            assert BandLayout(H, 16, n, i).rows == lib.partition_rows(H, n, i)
