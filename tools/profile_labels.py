"""A few label-driven fused losses (no CUDA graph) for ncu: python tools/profile_labels.py [cfg2|cfg3]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pea_b200 as P  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
lab = torch.tensor(bench.make_labels(w, 0).astype(np.int32)).to(dev)
g = torch.Generator(device="cpu").manual_seed(0)
emb = torch.randn((w["B"], 16) + tuple(w["spatial"]), generator=g).to(dev).requires_grad_(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
crit = P.WeightedMSE()
for _ in range(3):
    flush.zero_()
    if w["nd"] == 2:
        loss, _, _ = P.embedding_loss_labels(emb, lab, crit, P.multi_offset(w["shifts"], 4))
    else:
        loss, _ = P.embedding_loss_norm5_labels(emb, lab, crit)
    loss.backward()
    emb.grad = None
torch.cuda.synchronize()
print("ok", name, float(loss.detach()))
