"""Randomised stress of the canonical 2-D kernels (one view, two views, label-driven) against the table-driven generic
kernel: small and odd shapes (H >= 32, W >= 64, W % 4 == 0), random shifts <= 12, both borders, every mask type.
usage: python tools/stress2d.py [n_cases] [seed]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pea_b200 as P  # noqa: E402
from pea_b200 import _ops as ops  # noqa: E402
from oracle import pea_oracle as O  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst = 0.0
for case in range(n_cases):
    B = int(rng.randint(1, 4))
    H = int(rng.choice([32, 33, 40, 47, 48, 63, 64, 65, 100, 129]))
    W = int(rng.choice([64, 68, 72, 96, 100, 128, 132, 200]))
    shifts = sorted(int(v) for v in rng.choice(np.arange(1, 13), 5, replace=False))
    offsets = P.multi_offset(shifts, 4)
    offs3 = ops.offsets3(offsets, 2)
    K = 10
    g = torch.Generator().manual_seed(case)
    emb = torch.randn((B, 16, 1, H, W), generator=g).cuda()
    par = torch.randn((B, 16, 1, H, W), generator=g).cuda()
    tgt = (torch.rand((B, K, 1, H, W), generator=g) > 0.5).float().cuda()
    wgt = (torch.rand((B, K, 1, H, W), generator=g) * 2 + 0.1).cuda()
    m8 = (torch.rand((B, K, 1, H, W), generator=g) > 0.2).to(torch.uint8).cuda()
    border = ops.BORDER_WRAP if rng.rand() < 0.6 else ops.BORDER_ZERO
    msk = [m8, m8.float() * 0.5 + 0.5, None][int(rng.randint(3))]
    partner = par if rng.rand() < 0.4 else None
    cap = int(rng.choice([0, 1, 2, 5]))
    if cap:
        P._lib.lib.pea_debug_configure(1, int(cap))  # PEA_KNOB_MAX_CTAS
    else:
        P._lib.lib.pea_debug_configure(1, 0)
    norm = np.full(K, float(B * W), np.float32)
    coef = np.linspace(0.5, 1.5, K).astype(np.float32)
    outs = []
    for impl in (ops.IMPL_GENERIC, ops.IMPL_AUTO):
        lo, a, gr, _ = ops.affs_loss_fwd_bwd(emb, partner, offs3, border, tgt, wgt, msk, norm, coef, impl=impl)
        outs.append((lo, a, gr))
    errs = [float((outs[1][i] - outs[0][i]).abs().max() / outs[0][i].abs().max()) for i in range(3)]
    # label-driven flavour against the two-step route (circular border, as embedding_loss uses it)
    labs = np.stack([O.voronoi_labels((H, W), 7, case * 10 + b, background_frac=0.3, dtype=np.uint8) for b in range(B)])
    lab_d = torch.tensor(labs.astype(np.int32)).cuda()
    ign = bool(rng.rand() < 0.5)
    e1 = emb[:, :, 0].clone().requires_grad_(True)
    l1, a1, _ = P.embedding_loss_labels(e1, lab_d, P.WeightedMSE(), offsets, ignore=ign)
    l1.backward()
    t2, m2, w2 = P.make_targets_2d(lab_d, offsets, ignore=ign, padding=True)
    P._lib.lib.pea_debug_configure(0, 1)  # PEA_KNOB_NO_CANON
    e2 = emb[:, :, 0].clone().requires_grad_(True)
    l2, a2, _ = P.embedding_loss(e2, t2, w2, m2, P.WeightedMSE(), offsets)
    l2.backward()
    P._lib.lib.pea_debug_configure(0, 0)
    errs += [abs(float(l1) - float(l2)) / abs(float(l2)), float((a1 - a2).abs().max() / a2.abs().max()),
             float((e1.grad - e2.grad).abs().max() / e2.grad.abs().max())]
    worst = max(worst, max(errs))
    flag = "" if max(errs) <= 1e-5 and all(np.isfinite(errs)) else "   <-- FAIL"
    print(f"case {case}: B={B} H={H} W={W} shifts={shifts} border={border} mask={None if msk is None else msk.dtype} "
          f"two_view={partner is not None} cap={cap} ignore={ign}: " + " ".join(f"{e:.1e}" for e in errs) + flag, flush=True)
P._lib.lib.pea_debug_configure(1, 0)
print("worst", worst, "OK" if worst <= 1e-5 else "FAIL")
