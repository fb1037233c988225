"""BASELINE config 5 in miniature, under torchrun (one rank per GPU): a small conv backbone + the fused PEA loss under
DistributedDataParallel.  The loss kernels contain no collective and normalise with the LOCAL batch size, so DDP's mean
all-reduce of the backbone gradients must reproduce the single-process gradient of the concatenated global batch
(SURVEY.md 8e).  Prints one OK line per rank with the measured step time.
usage: torchrun --nproc-per-node N tools/ddp_step_check.py [2d|3d]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pea_b200 as P  # noqa: E402
from oracle import pea_oracle as O  # noqa: E402  (synthetic labels only)

mode = sys.argv[1] if len(sys.argv) > 1 else "3d"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
Bl = 2                                      # per-rank batch
torch.manual_seed(0)
if mode == "3d":
    net = nn.Sequential(nn.Conv3d(1, 8, 3, padding=1), nn.ELU(), nn.Conv3d(8, 16, 3, padding=1)).to(dev)
    shape = (18, 64, 64)
else:
    net = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ELU(), nn.Conv2d(8, 16, 3, padding=1)).to(dev)
    shape = (128, 128)
    offsets = P.multi_offset([1, 3, 5, 9, 11], 4)
g = torch.Generator().manual_seed(1)
imgs = torch.randn((world * Bl, 1 if mode == "3d" else 3) + shape, generator=g).to(dev)          # the GLOBAL batch, same on every rank
labs = np.stack([O.voronoi_labels(shape, 12, 50 + i, z_scale=4.0 if mode == "3d" else 1.0,
                                  dtype=np.uint16 if mode == "3d" else np.uint8) for i in range(world * Bl)])
lab_t = torch.tensor(labs.astype(np.int32) if mode == "3d" else labs).to(dev)
crit = P.WeightedMSE()


def loss_of(model, x, lab):
    emb = model(x)
    if mode == "3d":
        tgt, wgt = P.make_targets_3d(lab)
        return P.embedding_loss_norm5(emb, tgt, wgt, crit)[0]
    tgt, msk, wgt = P.make_targets_2d(lab, offsets)
    return P.embedding_loss(emb, tgt, wgt, msk, crit, offsets)[0]


# single-process reference on the whole global batch (every rank computes it redundantly)
ref = [p.detach().clone() for p in net.parameters()]
loss_ref = loss_of(net, imgs, lab_t)
loss_ref.backward()
gref = [p.grad.detach().clone() for p in net.parameters()]
net.zero_grad(set_to_none=True)

ddp = nn.parallel.DistributedDataParallel(net, device_ids=[local])
sl = slice(rank * Bl, (rank + 1) * Bl)
for it in range(3):
    ddp.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss = loss_of(ddp, imgs[sl], lab_t[sl])
    loss.backward()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
worst = 0.0
for p, gr in zip(net.parameters(), gref):
    worst = max(worst, float((p.grad - gr).abs().max() / gr.abs().max().clamp_min(1e-20)))
# tolerance: the backbone's cuDNN convolution backward sums a batch of 2 in a different order than a batch of 2 * world
# (measured 3e-4 relative in 3-D, 8e-5 in 2-D); the loss kernels themselves are held to 1e-5 by tests/test_gpu_parity.py
lsum = loss.detach().clone()
dist.all_reduce(lsum)
ok = worst < 2e-3 and abs(float(lsum) / world - float(loss_ref)) < 1e-4 * abs(float(loss_ref))
print(f"ddp_step_check {mode} rank {rank}/{world}: grad rel err vs global batch {worst:.2e}, mean loss {float(lsum) / world:.6f} "
      f"vs {float(loss_ref):.6f}, step {dt * 1e3:.2f} ms -> {'OK' if ok else 'FAIL'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
