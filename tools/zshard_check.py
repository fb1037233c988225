"""Run under torchrun (one rank per GPU): z-sharded affinities with the real NCCL halo exchange must equal, bit for
bit, the same slab cut out of the whole-volume result computed locally.  Prints one OK line per rank."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pea_b200 as P  # noqa: E402
from pea_b200.zshard import ZShard, sharded_affinities  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
D, H, W = 37, 96, 128
g = torch.Generator(device="cpu").manual_seed(7)
vol = torch.randn((1, 16, D, H, W), generator=g).to(dev)   # identical on every rank
whole = P.inf_embedding_loss_norm5(vol)
sh = ZShard(1, 16, D, H, W, rank, world, dev)
sh.own.copy_(vol[:, :, sh.z0:sh.z1])
if sh.halo:
    sh.buf[:, :, :sh.halo].fill_(float("nan"))
for it in range(3):
    out = sharded_affinities(sh)
    ok = torch.equal(out, whole[:, :, sh.z0:sh.z1])
    assert ok, f"rank {rank} iteration {it}: slab differs from the whole-volume result"
# fused halo through symmetric memory: no exchange step, the kernel reads the peer's slab tail in place
from pea_b200.zshard import PeerZShard, fused_halo_affinities, overlapped_halo_affinities  # noqa: E402
try:
    ps = PeerZShard(1, 16, D, H, W, rank, world, dev)
except Exception as ex:  # noqa: BLE001
    ps = None
    print(f"rank {rank}: symmetric memory unavailable ({type(ex).__name__}: {ex})", flush=True)
if ps is not None:
    # the slab is REWRITTEN every iteration (a real inference loop): publish / release inside affinities() must keep a
    # neighbour's in-flight reads of the old tail apart from the next write
    for it in range(6):
        g2 = torch.Generator(device="cpu").manual_seed(100 + it)
        vol_it = torch.randn((1, 16, D, H, W), generator=g2).to(dev)
        whole_it = P.inf_embedding_loss_norm5(vol_it)
        ps.own.copy_(vol_it[:, :, ps.z0:ps.z1])
        out = ps.affinities(overlapped=(it % 2 == 1))
        assert torch.equal(out, whole_it[:, :, ps.z0:ps.z1]), f"rank {rank} iteration {it}: peer-halo slab differs"
    print(f"zshard_check fused-halo OK rank {rank}/{world}", flush=True)
torch.cuda.synchronize()
dist.barrier()
print(f"zshard_check OK rank {rank}/{world} slab [{sh.z0},{sh.z1}) halo {sh.halo_bytes()} B", flush=True)
dist.destroy_process_group()
