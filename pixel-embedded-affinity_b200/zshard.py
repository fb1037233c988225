"""z-sharded full-volume affinity inference (BASELINE config 4).

A (B,C,D,H,W) embedding volume is split along z into one slab per rank.  All z offsets of the AC3/AC4 set point
backwards (-1..-4) and the y/x offsets stay inside a slice, so rank r needs exactly the LAST FOUR slices of rank r-1
and nothing else: one one-directional nearest-neighbour exchange (NCCL send/recv over NVLink), no other collective.
Rank 0 has no halo -- there the reference's zero border applies.

The slab buffer is allocated with room for the halo in front, so the halo is received in place and the kernel reads
the (halo + slab) range through a z-range view (channel-plane stride passed down the C ABI): no concatenation copy.
The exchange overlaps the slab interior: slices >= 4 of the slab do not depend on the halo and are computed first.

The reference never runs whole volumes through this op (it blends overlapping patches on the host,
scripts_ac3ac4/data/provider_valid.py:306-348); the per-voxel arithmetic is that of
``inf_embedding_loss_norm5`` (scripts_ac3ac4/loss/loss_embedding_mse.py:212-234).
"""
import torch
import torch.distributed as dist

HALO = 4  # largest z shift of the 12-offset set


def shard_ranges(D, world):
    """[z0, z1) of every rank: sizes differ by at most one, larger slabs first (100 over 8 -> 13,13,13,13,12,12,12,12)."""
    base, extra = divmod(D, world)
    out, z = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((z, z + n))
        z += n
    return out


class ZShard:
    """One rank's slab with the halo slot in front.  ``own`` is the view the producer (the backbone) writes into."""

    def __init__(self, B, C, D_total, H, W, rank, world, device, dtype=torch.float32):
        self.rank, self.world = rank, world
        self.z0, self.z1 = shard_ranges(D_total, world)[rank]
        self.depth = self.z1 - self.z0
        if world > 1 and self.depth < HALO:
            raise ValueError(f"slab of {self.depth} slices is thinner than the {HALO}-slice halo: use fewer ranks")
        self.halo = HALO if rank > 0 else 0
        self.buf = torch.empty((B, C, self.halo + self.depth, H, W), device=device, dtype=dtype)
        self.own = self.buf[:, :, self.halo:]

    def halo_bytes(self):
        B, C, _, H, W = self.buf.shape
        return B * C * self.halo * H * W * self.buf.element_size()


def start_halo_exchange(shard, group=None):
    """Post the sends/receives of the one-directional halo: last HALO slices -> next rank, in place.  Returns requests."""
    if shard.world == 1:
        return []
    B, C = shard.buf.shape[:2]
    ops = []
    for b in range(B):
        for c in range(C):
            if shard.rank + 1 < shard.world:   # every (b, c) plane's last slices are one contiguous chunk
                ops.append(dist.P2POp(dist.isend, shard.own[b, c, shard.depth - HALO:], shard.rank + 1, group))
            if shard.rank > 0:
                ops.append(dist.P2POp(dist.irecv, shard.buf[b, c, :HALO], shard.rank - 1, group))
    return dist.batch_isend_irecv(ops) if ops else []


def _cuda_compute(emb_view, z_skip, out_view, halo=None):
    from .loss_embedding_mse_3d import _inf3d, SHIFTS_NORM5
    _inf3d(emb_view, SHIFTS_NORM5, z_skip=z_skip, out=out_view, halo=halo)


def sharded_affinities(shard, out=None, group=None, compute=None):
    """12-channel affinities of this rank's slab, (B,12,depth,H,W).  ``compute(emb_view, z_skip, out_view)`` defaults to
    the CUDA kernel; tests on CPU inject the oracle to exercise the sharding logic with the gloo backend."""
    compute = compute or _cuda_compute
    B, C, _, H, W = shard.buf.shape
    if out is None:
        out = torch.empty((B, 12, shard.depth, H, W), device=shard.buf.device, dtype=torch.float32)
    reqs = start_halo_exchange(shard, group)
    head = min(HALO, shard.depth)
    # slab interior first: it only needs the slab itself (its first 4 slices warm the z ring up)
    if shard.depth > head:
        compute(shard.own, head, out[:, :, head:])
    for r in reqs:
        r.wait()
    # first slices of the slab: need the halo (rank > 0) or the zero border (rank 0)
    compute(shard.buf[:, :, :shard.halo + head], shard.halo, out[:, :, :head])
    return out


# ------------------------------------------------------------------------------------------------------------------
# Fused halo: no exchange step at all.  The slabs live in symmetric memory (every rank's buffer is mapped into every
# other rank's address space over NVLink / NVSwitch); rank r's kernel reads the last four slices of rank r-1 IN PLACE
# through the `halo` pointer of pea_affs_fwd (include/pea_b200.h) while it warms its z ring up.  268 MB per boundary
# cross NVLink exactly once, inside the compute kernel, overlapped with its own HBM traffic; nothing is staged.
class PeerZShard:
    """One rank's slab in symmetric memory.  ``own`` is what the producer writes; ``prev_tail`` is the (B,C,4,H,W) view
    of the previous rank's last four slices in PEER memory (None on rank 0)."""

    def __init__(self, B, C, D_total, H, W, rank, world, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.rank, self.world = rank, world
        ranges = shard_ranges(D_total, world)
        self.z0, self.z1 = ranges[rank]
        self.depth = self.z1 - self.z0
        if world > 1 and min(b - a for a, b in ranges) < HALO:
            raise ValueError(f"a slab is thinner than the {HALO}-slice halo: use fewer ranks")
        cap = max(b - a for a, b in ranges)            # symmetric allocations have the same size on every rank
        self.shape = (B, C, cap, H, W)
        self.buf = symm_mem.empty(self.shape, dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.buf, group=(group or dist.group.WORLD))
        self.own = self.buf[:, :, :self.depth]
        self.halo = HALO if rank > 0 else 0
        self.prev_tail = None
        if rank > 0:
            pd = ranges[rank - 1][1] - ranges[rank - 1][0]
            peer = self.handle.get_buffer(rank - 1, self.shape, torch.float32)
            self.prev_tail = peer[:, :, pd - HALO:pd]

    def halo_bytes(self):
        B, C, _, H, W = self.shape
        return B * C * self.halo * H * W * 4

    def staging(self):
        """Local landing buffer + side stream for the overlapped variant (allocated on first use)."""
        if getattr(self, "_stage", None) is None:
            B, C, _, H, W = self.shape
            self._stage = torch.empty((B, C, HALO, H, W), device=self.buf.device, dtype=torch.float32)
            self._side = torch.cuda.Stream(device=self.buf.device)
        return self._stage, self._side

    # ---- the two-barrier contract of a slab that is rewritten every step -----------------------------------------
    #   write own  ->  publish()  ->  kernel (reads prev_tail in place)  ->  release()  ->  next write of own
    # publish() orders every producer's writes before the neighbours' reads; release() orders every rank's reads of its
    # neighbour's tail before that neighbour's NEXT write.  Without the second barrier rank r-1 may overwrite its slab
    # while rank r's kernel is still reading the last four slices over NVLink (a cross-GPU write-after-read race that
    # corrupts slab-boundary affinities silently).  ``affinities()`` issues both.
    def publish(self):
        """Order the producer's writes of ``own`` before the neighbours' reads (device-side barrier on the current
        stream; no host synchronisation)."""
        if self.world > 1:
            self.handle.barrier(channel=0)

    def release(self):
        """Order this step's reads of the neighbours' slab tails (the kernels enqueued since ``publish()``) before anyone's
        next write of ``own`` (device-side barrier on the current stream, second channel)."""
        if self.world > 1:
            self.handle.barrier(channel=1)

    def affinities(self, out=None, overlapped=False, compute=None, staged=None):
        """publish() -> 12-channel affinities of this rank's slab -> release().  After it returns (stream-ordered)
        ``own`` may be rewritten for the next volume.  Default: the kernel reads the tail from peer memory in place
        (``fused_halo_affinities``).  ``staged=True``: a copy engine pulls the peer's tail into a local buffer band by band
        while the ONE compute launch runs (``staged_halo_affinities``) -- same bits, measured no faster on NVLink 5
        (0.556 - 0.569 ms against 0.555 ms per 13-slice slab at 1024 x 1024, 2 x B200), kept for slower links."""
        staged = bool(staged) and compute is None and not overlapped
        self.publish()
        if overlapped:
            out = overlapped_halo_affinities(self, out=out, compute=compute)
        elif staged and self.prev_tail is not None:
            out = staged_halo_affinities(self, out=out)
        else:
            out = fused_halo_affinities(self.own, self.prev_tail, out=out, compute=compute)
        self.release()
        return out


def fused_halo_affinities(emb_own, prev_tail, out=None, compute=None):
    """12-channel affinities of a slab in ONE launch: ``prev_tail`` (B,C,<=4,H,W; None for the first slab) is read in
    place -- typically from the peer GPU.  Bit-identical to the same slices of the whole-volume result.
    With peer memory the caller owns the ordering: ``PeerZShard.publish()`` before, ``PeerZShard.release()`` after
    (or call ``PeerZShard.affinities()``, which does both)."""
    compute = compute or _cuda_compute
    B, C, depth, H, W = emb_own.shape
    if out is None:
        out = torch.empty((B, 12, depth, H, W), device=emb_own.device, dtype=torch.float32)
    compute(emb_own, 0, out, prev_tail)
    return out


# Staged halo (round 2, ABI 7): the third way to get the four slices across.  `fused_halo_affinities` makes every CTA
# pull its columns of the tail over NVLink itself (one launch); `overlapped_halo_affinities` uses the copy engine but
# needs two launches and warms the ring up twice.  Here the copy engine pulls the tail in BANDS OF IMAGE ROWS, in the
# order the kernel's tile rows are scheduled, each band followed by a stream-ordered flag write; the single compute
# launch waits per tile row for its band's flag (bounded: a band later than 40 us is read from the peer instead, so the
# kernel cannot hang on a copy that does not run) and reads local memory.  The first STAGED_ROW0 rows -- whose CTAs are
# resident before any band can have landed -- read the peer directly as before.
# Measured (tools/zshard_staged.py, 2 x B200, 13-slice slabs of 1024 x 1024): kernel without a halo 0.472 ms, in-place
# halo 0.555 ms per step, staged 0.556 (row0 160, bands of 128 rows) - 0.569 ms (row0 96, bands of 64) - 0.620 (bands of
# 256): no gain.  The same kernel with the tail in LOCAL memory (one GPU) takes 0.522 ms: of the 83 us a halo costs, 56
# are the ring warm-up itself (268 MB more to read, a second warm-up for the CTAs that visit the halo slices last) and
# 15 the two barriers; NVLink exposure is the remaining ~15 us, and the copy engine's own traffic gives those back.
STAGED_ROW0 = 160        # first waves of CTAs (296 resident CTAs = 74 rows of 32-column tiles at W = 1024) read in place
STAGED_BAND_ROWS = 128


def staged_halo_affinities(shard, out=None, row0=None, band_rows=None):
    """12-channel affinities of ``shard.own`` with the previous rank's tail pulled by a copy engine behind the running
    kernel (pea_zshard_stage_halo + pea_zshard_affs_fwd, include/pea_b200.h).  Bit-identical to
    ``fused_halo_affinities(shard.own, shard.prev_tail)``; call between ``publish()`` and ``release()``."""
    from ._lib import lib, check
    from ._ops import _ptr, _stream
    B, C, depth, H, W = shard.own.shape
    dev = shard.own.device
    if out is None:
        out = torch.empty((B, 12, depth, H, W), device=dev, dtype=torch.float32)
    row0 = min(STAGED_ROW0 if row0 is None else row0, H) // 8 * 8
    band_rows = STAGED_BAND_ROWS if band_rows is None else band_rows
    stage, side = shard.staging()
    nb = max(1, -(-(H - row0) // band_rows))
    if getattr(shard, "_flags", None) is None or shard._flags.numel() < nb:
        shard._flags = torch.zeros(nb, device=dev, dtype=torch.int32)
        shard._epoch = 0
    shard._epoch = (shard._epoch + 1) & 0xFFFFFFFF   # full 32-bit wrap: the kernel compares the signed distance flag - epoch
    tail = shard.prev_tail
    ts, os_, es = tail.stride(), out.stride(), shard.own.stride()
    cur = torch.cuda.current_stream(dev)
    side.wait_stream(cur)            # behind publish(): the producer's tail is complete; behind the previous step's kernel
    with torch.cuda.device(dev):
        if row0 < H:
            check("pea_zshard_stage_halo",
                  lib.pea_zshard_stage_halo(_ptr(tail), int(ts[1]), _ptr(stage), B, C, int(tail.shape[2]), H, W, row0, band_rows,
                                            _ptr(shard._flags), shard._epoch, side.cuda_stream))
        check("pea_zshard_affs_fwd",
              lib.pea_zshard_affs_fwd(_ptr(shard.own), B, C, depth, H, W, int(es[1]), int(os_[1]), _ptr(tail),
                                      int(tail.shape[2]), int(ts[1]), _ptr(stage), _ptr(shard._flags), shard._epoch, row0,
                                      band_rows, 0, _ptr(out), _stream()))
    cur.wait_stream(side)            # release() and the next publish() follow the last copy as well as the kernel
    return out


def overlapped_halo_affinities(shard, out=None, compute=None):
    """Same result as ``fused_halo_affinities(shard.own, shard.prev_tail)`` in two launches: the previous rank's tail is
    pulled over NVLink into a local buffer on a side stream (DMA, no SM time) WHILE the slab interior -- slices >= 4,
    which need no halo -- is computed; the first four slices follow, reading the local copy.  Measured on 8 B200 at
    100 x 1024 x 1024 (268 MB of halo against 13 slices per rank): 0.669 ms, against 0.618 ms for the single fused launch
    (whose CTAs overlap each other's NVLink reads) -- kept as the alternative for links slower than NVLink 5."""
    compute = compute or _cuda_compute
    B, C, depth, H, W = shard.own.shape
    if out is None:
        out = torch.empty((B, 12, depth, H, W), device=shard.own.device, dtype=torch.float32)
    if shard.prev_tail is None or depth <= HALO:
        compute(shard.own, 0, out, shard.prev_tail)
        return out
    stage, side = shard.staging()
    cur = torch.cuda.current_stream(shard.own.device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        for b in range(B):
            for c in range(C):   # one contiguous 4-slice chunk per (sample, channel)
                stage[b, c].copy_(shard.prev_tail[b, c], non_blocking=True)
    compute(shard.own, HALO, out[:, :, HALO:])          # interior: ring warmed up on the slab's own first slices
    cur.wait_stream(side)
    compute(shard.own[:, :, :HALO], 0, out[:, :, :HALO], stage)
    return out
