"""cfg4 only, under torchrun: z-sharded full-volume affinities (bench.time_zshard), one JSON line on rank 0."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pea_b200 as P  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=dev)
res = bench.time_zshard(P, dev, rank, world, int(sys.argv[1]) if len(sys.argv) > 1 else 10)
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
