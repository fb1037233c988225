"""Time seg -> loss + gradient at cfg3: two-step route (GT kernels + fused loss) vs the label-driven 3-D kernel."""
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pea_b200 as P  # noqa: E402

dev = torch.device("cuda", 0)
w = bench.WORKLOADS["cfg3"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
seg = torch.tensor(bench.make_labels(w, 0).astype(np.int32)).to(dev)
g = torch.Generator(device="cpu").manual_seed(0)
emb = torch.randn((w["B"], 16) + tuple(w["spatial"]), generator=g).to(dev)
crit = P.WeightedMSE()
e = emb.clone().requires_grad_(True)


def two_step():
    tgt, wgt = P.make_targets_3d(seg)
    loss, _ = P.embedding_loss_norm5(e, tgt, wgt, crit)
    loss.backward()
    e.grad = None
    return loss


def fused():
    loss, _ = P.embedding_loss_norm5_labels(e, seg, crit)
    loss.backward()
    e.grad = None
    return loss


for name, fn in (("two_step", two_step), ("fused", fused)):
    for _ in range(3):
        fn()
    per = []
    for _ in range(20):
        flush.zero_()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        loss = fn()
        t.record()
        torch.cuda.synchronize()
        per.append(s.elapsed_time(t))
    print(f"{name}: median {statistics.median(per) * 1e3:.1f} us  loss {float(loss.detach()):.6f}")
