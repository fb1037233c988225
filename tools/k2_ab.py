"""A/B harness for the 2-D fused kernels: AUTO (axis-aligned TMA kernel where eligible) against the table-driven kernel
on awkward shapes / masks / borders, then timing of cfg2 and cfg1.  usage: python tools/k2_ab.py [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pea_b200 as P  # noqa: E402
from pea_b200 import _ops as ops  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda", 0)


def run(impl, emb, par, offs3, border, tgt, wgt, msk, grad):
    K = offs3.shape[0]
    norm = np.full(K, float(emb.shape[0] * emb.shape[-1]), np.float32)
    coef = np.linspace(0.5, 1.5, K).astype(np.float32)
    lo, a, g, _ = ops.affs_loss_fwd_bwd(emb, par, offs3, border, tgt, wgt, msk, norm, coef, want_affs=True, want_grad=grad,
                                        impl=impl)
    return lo, a, g


worst = 0.0
cases = [(2, 40, 64, [1, 3, 5, 9, 11]), (1, 67, 96, [1, 3, 5, 9, 27]), (1, 130, 128, [1, 3, 5, 9, 11]), (1, 101, 200, [2, 12]),
         (2, 64, 72, [1, 3, 5, 9, 27]), (1, 150, 160, [1, 3, 5, 9, 27])]
for (B, H, W, shifts) in cases:
    offsets = P.multi_offset(shifts, 4)
    if len(shifts) == 2:
        offsets = [[s, 0] for s in shifts] + [[0, -s] for s in shifts] + [[0, 7]]   # mixed signs
    offs3 = ops.offsets3(offsets, 2)
    K = len(offsets)
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + H)
    emb = torch.randn((B, 16, 1, H, W), generator=g).to(dev)
    par = torch.randn((B, 16, 1, H, W), generator=g).to(dev)
    emb[0, :, 0, 3, 5] = 0.0
    tgt = (torch.rand((B, K, 1, H, W), generator=g) > 0.5).float().to(dev)
    wgt = (torch.rand((B, K, 1, H, W), generator=g) * 2 + 0.1).to(dev)
    m8 = (torch.rand((B, K, 1, H, W), generator=g) > 0.2).to(torch.uint8).to(dev)
    for border in (ops.BORDER_WRAP, ops.BORDER_ZERO):
        for msk in (m8, m8.float() * 0.5 + 0.5, None):
            for partner in (None, par):
                for grad in (True, False):
                    lg, ag, gg = run(ops.IMPL_GENERIC, emb, partner, offs3, border, tgt, wgt, msk, grad)
                    lt, at, gt = run(ops.IMPL_AUTO, emb, partner, offs3, border, tgt, wgt, msk, grad)
                    errs = [float((lt - lg).abs().max() / lg.abs().max()), float((at - ag).abs().max() / ag.abs().max())]
                    if grad:
                        errs.append(float((gt - gg).abs().max() / gg.abs().max()))
                    worst = max(worst, max(errs))
                    if not (max(errs) <= 1e-5) or not all(np.isfinite(errs)):
                        print(f"FAIL shape {(B, H, W)} shifts {shifts} border {border} mask {None if msk is None else msk.dtype} "
                              f"cross {partner is not None} grad {grad}: " + " ".join(f"{e:.2e}" for e in errs), flush=True)
    print(f"shape {(B, H, W)} shifts {shifts}: done, worst so far {worst:.2e}", flush=True)
print("worst", worst, "OK" if worst <= 1e-5 else "FAIL", flush=True)

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in ("cfg2", "cfg1", "cfg2x"):
    w = bench.WORKLOADS[name]
    r = bench.time_workload(P, w, steps, 5, dev, flush, False)
    per = sorted(r["per_ms"])
    ab, K = bench.algorithmic_bytes_per_step(w)
    ms = sum(per) / len(per)
    print(f"{name}: mean {ms * 1e3:.1f} us  min {per[0] * 1e3:.1f} us   {ab / (ms * 1e-3) / 1e9:.0f} GB/s = "
          f"{ab / (ms * 1e-3) / 1e9 / 6459.6 * 100:.1f}% of measured HBM peak   loss {r['loss']:.6f}", flush=True)
