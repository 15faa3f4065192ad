"""Sharding of one frame over N ranks (one process per GPU) and the buffers of its assembly.

The reference's only communication step is a gather of contiguous pixel ranges to rank 0 with blocking
MPI_Send/MPI_Recv (src/main-mpi.cxx:258-293, 854-881).  Here:

* `PeerLanes` -- the product path: L frames, images and barrier flags of all ranks in peer-mapped (symmetric) memory;
  the library's kernels store into them directly (srt_render_xray_blocks_device), nothing is gathered afterwards.
* `BandLayout` / `InterleavedLayout` -- the fallback without peer mappings (and what the gloo CPU tests cover): every
  rank renders its rows into a full-frame or slot buffer and a single in-place all_gather (NCCL over NVLink on the
  GPUs, gloo in the CPU tests) leaves the whole frame on every rank, so the +-4-pixel repair stencil
  (main-pthreads-lbuffer.cxx:329-404) can run after assembly and results do not depend on N.

Band partition rule: H / N rows each, the first H % N ranks get one more (pthreadWorkLoadAllocation,
main-pthreads-lbuffer.cxx:694-725; rows variant main-pthreads-rows.cxx:312-334).  all_gather needs equal counts, so
when N does not divide H the frame buffer is padded: bands are gathered in slots of ceil(H/N) rows and compacted
afterwards.
"""
import torch
import torch.distributed as dist


class BandLayout:
    def __init__(self, H, W, world, rank):
        self.H, self.W, self.world, self.rank = H, W, world, rank
        per, rem = divmod(H, world)
        self.starts = [i * per + min(i, rem) for i in range(world)]
        self.counts = [per + (1 if i < rem else 0) for i in range(world)]
        self.rows = (self.starts[rank], self.starts[rank] + self.counts[rank])
        self.even = rem == 0
        self.slot_rows = per + (0 if self.even else 1)

    def new_frame(self, dtype, device):
        """Buffer a rank renders into: the H*W frame (+ one slot of staging when bands are uneven)."""
        n = self.H * self.W + (0 if self.even else self.world * self.slot_rows * self.W)
        return torch.empty(n, dtype=dtype, device=device)

    def all_gather_(self, frame):
        """Assemble the full frame on every rank. Returns a view of the first H*W elements."""
        H, W = self.H, self.W
        out = frame[: H * W]
        if self.world == 1:
            return out
        if self.even:
            mine = out.view(self.world, -1)[self.rank]
            dist.all_gather_into_tensor(out, mine)                       # in place: send slot == my band
            return out
        slots = frame[H * W:].view(self.world, self.slot_rows * W)
        rb, re = self.rows
        send = slots[self.rank]
        send[: (re - rb) * W].copy_(out[rb * W: re * W])
        dist.all_gather_into_tensor(slots.view(-1), send.clone())
        for i in range(self.world):
            if i != self.rank and self.counts[i]:
                out[self.starts[i] * W: (self.starts[i] + self.counts[i]) * W].copy_(slots[i][: self.counts[i] * W])
        return out


class InterleavedLayout:
    """Block-cyclic sharding of the frame's 16-row blocks over the ranks (load balance: rows through the object
    take the exact path far more often than empty rows, so contiguous bands finish at very different times).

    Rank r renders blocks r, r + world, ... into its dense slot (`slot_rows` rows, see srt_interleave_slot_rows);
    one in-place all_gather of the slots leaves `slots[world][slot_rows][W]` on every rank; de-interleaving back
    to image order is srt_deinterleave_device on the GPU (`deinterleave_torch` is the same mapping in torch for
    the CPU tests of this host logic).
    """
    BLOCK = 16

    def __init__(self, H, W, world, rank):
        self.H, self.W, self.world, self.rank = H, W, world, rank
        self.n_blocks = (H + self.BLOCK - 1) // self.BLOCK
        self.slot_rows = ((self.n_blocks + world - 1) // world) * self.BLOCK

    def new_slots(self, dtype, device):
        return torch.zeros(self.world * self.slot_rows * self.W, dtype=dtype, device=device)

    def my_slot(self, slots):
        return slots.view(self.world, self.slot_rows * self.W)[self.rank]

    def rows_of(self, rank):
        """image rows rendered by `rank`, in slot order"""
        rows = []
        for b in range(rank, self.n_blocks, self.world):
            rows.extend(range(b * self.BLOCK, min((b + 1) * self.BLOCK, self.H)))
        return rows

    def all_gather_(self, slots):
        if self.world > 1:
            dist.all_gather_into_tensor(slots, self.my_slot(slots))          # in place: send slot == my slot
        return slots

    def deinterleave_torch(self, slots):
        out = torch.empty(self.H, self.W, dtype=slots.dtype, device=slots.device)
        s = slots.view(self.world, self.slot_rows, self.W)
        for r in range(self.world):
            rows = self.rows_of(r)
            if rows:
                out[torch.tensor(rows)] = s[r, : len(rows)]
        return out.view(-1)


class PeerLanes:
    """Frame buffers and barrier flags of a PIPELINE of sharded frames, in one symmetric allocation per rank.

    Frames are rendered round-robin on `lanes` independent streams (one srt context each), so that the short,
    latency-bound kernels of one frame (candidate stages, hit-list resolve, the device barriers, the repair pass)
    overlap the scan of the next.  A lane owns two sets of {L frame, image} used alternately and one flag array for
    srt_peer_barrier_device (two barriers per frame: after the L rows, after the repaired image rows).  Layout per
    lane, in 4-byte words: [L frame 0: H*W][L frame 1][image 0][image 1][flags: 64].
    """
    FLAG_WORDS = 64

    def __init__(self, H, W, lanes, group=None, use_multicast=True):
        import torch.distributed._symmetric_memory as symm_mem
        self.H, self.W, self.lanes = H, W, lanes
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.lane_words = 4 * H * W + self.FLAG_WORDS
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buf = symm_mem.empty(lanes * self.lane_words, dtype=torch.float32, device=dev)
        self.buf.zero_()                                      # flags start at epoch 0
        torch.cuda.synchronize()
        self.handle = symm_mem.rendezvous(self.buf, group)
        dist.barrier(group)                                   # every rank's flags are zero before anyone signals
        self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        mc = 0
        if use_multicast:
            try:
                mc = int(self.handle.multicast_ptr)
            except Exception:
                mc = 0
        self.multicast_ptr = mc

    def _frame_off(self, lane, s):
        return lane * self.lane_words + s * self.H * self.W              # words

    def frame(self, lane, s):
        """this rank's frame of set `s` of `lane` (a view of the symmetric buffer)"""
        o = self._frame_off(lane, s)
        return self.buf[o: o + self.H * self.W]

    def frame_ptrs(self, lane, s):
        """-> (frame pointers for srt_render_*_blocks_device, is_multicast): the multicast mapping when the fabric
        offers one, else own frame first, then each peer's"""
        o = 4 * self._frame_off(lane, s)
        if self.multicast_ptr:
            return [self.multicast_ptr + o], True
        return [self.peer_ptrs[self.rank] + o] + [p + o for r, p in enumerate(self.peer_ptrs) if r != self.rank], False

    def unicast_ptrs(self, lane, s):
        """own buffer first, then every peer's in rank order (what SRT_FLAG_L_HALO_ONLY expects)"""
        o = 4 * self._frame_off(lane, s)
        return [self.peer_ptrs[self.rank] + o] + [p + o for r, p in enumerate(self.peer_ptrs) if r != self.rank]

    def own_frame_ptr(self, lane, s):
        return self.peer_ptrs[self.rank] + 4 * self._frame_off(lane, s)

    def image(self, lane, s):
        """this rank's repaired image of set `s` of `lane`"""
        return self.frame(lane, 2 + s)

    def image_ptrs(self, lane, s):
        return self.frame_ptrs(lane, 2 + s)

    def flag_ptrs(self, lane):
        """flag array of every rank, in rank order, as mapped here"""
        o = 4 * (lane * self.lane_words + 4 * self.H * self.W)
        return [p + o for p in self.peer_ptrs]

    def poison(self, value=float("nan")):
        """overwrite every L frame and image of this rank (not the flags)"""
        for lane in range(self.lanes):
            for s in range(4):
                self.frame(lane, s).fill_(value)
