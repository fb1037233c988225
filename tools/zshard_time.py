"""Time the z-sharded forward kernel at a given volume depth under torchrun (one rank per GPU) and check every rank's
slab bit for bit first.  usage: torchrun ... tools/zshard_time.py [D] [steps]   (D = 26 at 2 GPUs mimics the 13-slice
slabs of the 100-slice volume on 8 GPUs: rank 1 reads a 268 MB halo, rank 0 none)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pea_b200 as P  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
D = int(sys.argv[1]) if len(sys.argv) > 1 else 26
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
r = bench.time_zshard(P, dev, rank, world, steps, shape=(D, 1024, 1024))
if rank == 0:
    print(f"zshard D={D} world={world}: {r['ms_per_step']:.4f} ms  halo_check={r['halo_check']}  transport={r['halo_transport']}", flush=True)
dist.barrier()
dist.destroy_process_group()
