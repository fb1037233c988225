"""A/B harness for the fused 3-D kernel: correctness of the tiled kernel against the table-driven one on awkward
shapes, then timing of one bench workload (CUDA-graph replay, L2 flushed).  Select the build with PEA_B200_LIB.
usage: python tools/k4_ab.py [cfg3|cfg3x|cfg2|cfg1] [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pea_b200 as P  # noqa: E402
from pea_b200 import _ops as ops  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda", 0)


def run(impl, emb, tgt, wgt, grad=True):
    ops.set_default_impl(impl)
    try:
        e = emb.clone().requires_grad_(grad)
        if grad:
            l, a = P.embedding_loss_norm5(e, tgt, wgt, P.WeightedMSE(), affs0_weight=1.5)
            l.backward()
            return l.detach(), a, e.grad
        with torch.no_grad():
            l, a = P.embedding_loss_norm5(e, tgt, wgt, P.WeightedMSE(), affs0_weight=1.5)
        return l, a, None
    finally:
        ops.set_default_impl(ops.IMPL_AUTO)


worst = 0.0
for (B, D, H, W) in [(1, 5, 28, 32), (2, 7, 37, 44), (1, 18, 64, 96), (1, 6, 30, 28), (1, 5, 40, 36), (2, 11, 50, 68), (1, 9, 30, 40)]:
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + D * 100 + H)
    emb = torch.randn((B, 16, D, H, W), generator=g).to(dev)
    emb[0, :, 0, 3, 5] = 0.0
    tgt = (torch.rand((B, 12, D, H, W), generator=g) > 0.5).float().to(dev)
    wgt = (torch.rand((B, 12, D, H, W), generator=g) * 2 + 0.1).to(dev)
    for grad in (True, False):
        lg, ag, gg = run(ops.IMPL_GENERIC, emb, tgt, wgt, grad)
        lt, at, gt = run(ops.IMPL_TILED, emb, tgt, wgt, grad)
        errs = [abs(float(lt) - float(lg)) / abs(float(lg)), float((at - ag).abs().max() / ag.abs().max())]
        if grad:
            errs.append(float((gt - gg).abs().max() / gg.abs().max()))
        worst = max(worst, max(errs))
        flag = "" if max(errs) <= 1e-5 and all(np.isfinite(errs)) else "   <-- FAIL"
        print(f"shape {(B, D, H, W)} grad={grad}: rel err loss/affs/grad = " + " ".join(f"{e:.2e}" for e in errs) + flag, flush=True)
print("worst", worst, "OK" if worst <= 1e-5 else "FAIL", flush=True)

w = bench.WORKLOADS[name]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
r = bench.time_workload(P, w, steps, 5, dev, flush, False)
per = sorted(r["per_ms"])
ab, K = bench.algorithmic_bytes_per_step(w)
ms = sum(per) / len(per)
print(f"{name}: mean {ms * 1e3:.1f} us  median {per[len(per) // 2] * 1e3:.1f} us  min {per[0] * 1e3:.1f} us   "
      f"{ab / (ms * 1e-3) / 1e9:.0f} GB/s = {ab / (ms * 1e-3) / 1e9 / 6459.6 * 100:.1f}% of measured HBM peak   loss {r['loss']:.6f}",
      flush=True)
