"""Where a z-shard step's time goes (torchrun, one rank per GPU): the full publish -> kernel -> release step, the kernel
alone (with and without the peer halo), and the two device-side barriers alone.  usage: torchrun ... tools/zshard_parts.py [D]"""
import os, sys, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pea_b200 as P
from pea_b200.zshard import PeerZShard, fused_halo_affinities
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
D = int(sys.argv[1]) if len(sys.argv) > 1 else 13 * world
sh = PeerZShard(1, 16, D, 1024, 1024, rank, world, dev)
sh.own.normal_()
out = torch.empty((1, 12, sh.depth, 1024, 1024), device=dev)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev, dtype=torch.float64)
    lo = t.clone(); dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return float(t), float(lo)
full = timeit(lambda: sh.affinities(out=out))
kern = timeit(lambda: fused_halo_affinities(sh.own, sh.prev_tail, out=out))
nohalo = timeit(lambda: fused_halo_affinities(sh.own, None, out=out))
def bars():
    sh.publish(); sh.release()
bar = timeit(bars)
if rank == 0:
    print(f"world={world} D={D} depth={sh.depth}: full step max {full[0]:.4f} (min {full[1]:.4f}) ms | kernel with halo max {kern[0]:.4f} min {kern[1]:.4f} | "
          f"kernel without halo {nohalo[0]:.4f} | two barriers {bar[0]*1e3:.1f} us", flush=True)
dist.barrier(); dist.destroy_process_group()
