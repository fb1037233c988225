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


def _cuda_compute(emb_view, z_skip, out_view):
    from .loss_embedding_mse_3d import _inf3d, SHIFTS_NORM5
    _inf3d(emb_view, SHIFTS_NORM5, z_skip=z_skip, out=out_view)


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
