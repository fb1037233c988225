"""Run the batch GT generation of one bench workload a few times (for ncu).  usage: python tools/profile_gt.py [cfg2|cfg3]"""
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
lab = bench.make_labels(w, 0)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
if w["nd"] == 2:
    t = torch.tensor(lab).to(dev)
    offsets = P.multi_offset(w["shifts"], 4)
    fn = lambda: P.make_targets_2d(t, offsets)
else:
    t = torch.tensor(lab.astype(np.int32)).to(dev)
    fn = lambda: P.make_targets_3d(t)
for _ in range(3):
    flush.zero_()
    fn()
torch.cuda.synchronize()
print("ok", name)
