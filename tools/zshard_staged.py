"""Staged halo (pea_zshard_stage_halo + pea_zshard_affs_fwd) against the in-place peer read, per slab.
  one GPU  : python tools/zshard_staged.py [depth]          -- the "peer" tail is a second local buffer (HBM load of the
             pull is the same as on a real box: every GPU serves 268 MB to its neighbour and lands 268 MB)
  torchrun : ... tools/zshard_staged.py [D_total]           -- real peers, PeerZShard, max over ranks
Prints kernel-only and full-step (publish -> ... -> release) times for: no halo, in-place halo, staged halo with a few
(row0, band_rows) settings; checks staged == in-place bit for bit first."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pea_b200 as P  # noqa: F401,E402
from pea_b200 import zshard as Z  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
H = W = 1024
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    D = int(sys.argv[1]) if len(sys.argv) > 1 else 13 * world
    sh = Z.PeerZShard(1, 16, D, H, W, rank, world, dev)
    sh.own.normal_()
else:
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 13

    class Emu:
        def __init__(self):
            self.buf = torch.randn(1, 16, depth, H, W, device=dev)
            self.own = self.buf
            self.peer = torch.randn(1, 16, depth, H, W, device=dev)
            self.prev_tail = self.peer[:, :, depth - 4:]
            self.depth, self.world, self._stage = depth, 1, None
            self.shape = tuple(self.buf.shape)

        staging = Z.PeerZShard.staging

        def publish(self): pass

        def release(self): pass
    sh = Emu()
out = torch.empty((1, 12, sh.depth, H, W), device=dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def step(fn):
    def f():
        sh.publish(); fn(); sh.release()
    return f


res = {}
res["no halo (kernel)"] = timeit(lambda: Z.fused_halo_affinities(sh.own, None, out=out))
if sh.prev_tail is not None or world > 1:
    sh.publish()
    ref = Z.fused_halo_affinities(sh.own, sh.prev_tail).clone() if sh.prev_tail is not None else None
    sh.release()
    res["in-place halo, full step"] = timeit(step(lambda: Z.fused_halo_affinities(sh.own, sh.prev_tail, out=out)))
    from pea_b200._lib import lib as _lib
    for waves in (-1, 1, 2, 1 << 20):   # PEA_KNOB_K3_ROT_WAVES: which waves of CTAs visit the halo slices last
        assert _lib.pea_debug_configure(4, waves) == 0
        if sh.prev_tail is not None:
            sh.publish()
            assert torch.equal(Z.fused_halo_affinities(sh.own, sh.prev_tail), ref), f"rot waves {waves}"
            sh.release()
        else:
            sh.publish(); sh.release()
        res[f"in-place halo, rot waves {waves}, full step"] = timeit(step(lambda: Z.fused_halo_affinities(sh.own, sh.prev_tail, out=out)))
    _lib.pea_debug_configure(4, 0)
    for row0, band in ((160, 128),) if os.environ.get("PEA_STAGED_SWEEP", "0") != "1" else ((96, 64), (96, 128), (160, 128), (96, 256)):
        sh.publish()   # collective: every rank, also the one without a halo
        got = Z.staged_halo_affinities(sh, row0=row0, band_rows=band) if sh.prev_tail is not None else None
        sh.release()
        torch.cuda.synchronize()
        assert got is None or torch.equal(got, ref), f"staged ({row0}, {band}) differs from the in-place result"
        del got
        f = (lambda r=row0, b=band: Z.staged_halo_affinities(sh, out=out, row0=r, band_rows=b)) if sh.prev_tail is not None \
            else (lambda: Z.fused_halo_affinities(sh.own, None, out=out))
        res[f"staged row0={row0} band={band}, full step"] = timeit(step(f))
if rank == 0:
    print(f"world={world} depth={sh.depth} {H}x{W}")
    for k, v in res.items():
        print(f"  {k:45s} {v:.4f} ms", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
