"""Run a few fused fwd+bwd steps of one bench workload (no CUDA graph) so that ncu can attribute kernels.
usage: python tools/profile_step.py [cfg3|cfg2|cfg1] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pea_b200 as P  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
lab, emb, tgt, wgt, msk, offsets = bench.device_inputs(P, w, 0, dev)
crit = P.WeightedMSE()
e = emb.clone().requires_grad_(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(steps):
    flush.zero_()
    loss = bench.our_step(P, w, e, tgt, wgt, msk, offsets, crit)
    e.grad = None
torch.cuda.synchronize()
print("ok", name, float(loss.detach()))
