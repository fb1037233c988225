"""Randomised stress of the kernels added in the second session of round 2 (run by hand on a GPU box, not part of pytest):
  * forward-only 2-D kernel (pea_affs2d_fwd.cu) against the table-driven kernel: random B, H, W (W % 4 == 0), 1 - 6 random
    shifts up to 28, both borders, relu on / off, random grid caps;
  * batch GT in one launch against the two-launch route and the oracle: random shapes, label dtypes, offset tables
    (including positive and mixed-sign offsets), ignore / padding flags, 2-D and 3-D;
  * the work-list kernel as fallback of pea_affs_loss_fwd_bwd / pea_affs_bwd against the table-driven kernel.
usage: python tools/stress_r2b.py [n_cases] [seed]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pea_b200 as P  # noqa: E402
from pea_b200 import _ops as ops  # noqa: E402
from pea_b200._lib import lib  # noqa: E402
from oracle import pea_oracle as O  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)


def knob(k, v):
    assert lib.pea_debug_configure(k, int(v)) == 0


def worst(ours, ref):
    ours, ref = ours.double().cpu(), ref.double().cpu()
    scale = max(float(ref.abs().max()), 1e-30)
    return float(((ours - ref).abs() / (1e-5 * ref.abs() + 1e-6 * scale)).max())


bad = 0
for case in range(n_cases):
    # ---- forward-only 2-D
    B, H, W = int(rng.randint(1, 4)), int(rng.randint(32, 200)), 4 * int(rng.randint(16, 60))
    ns = int(rng.randint(1, 7))
    shifts = sorted(int(s) for s in rng.choice(np.arange(1, 29), size=ns, replace=False))
    offs3 = ops.offsets3(O.multi_offset(shifts, neighbor=4), 2)
    emb = torch.randn((B, 16, H, W), generator=torch.Generator().manual_seed(case)).cuda()
    border = ops.BORDER_WRAP if rng.rand() < 0.6 else ops.BORDER_ZERO
    relu = bool(rng.rand() < 0.5)
    knob(1, int(rng.choice([0, 1, 3, 17])))
    a = ops.affs_fwd(emb, None, offs3, border, relu=relu)
    knob(1, 0)
    g = ops.affs_fwd(emb, None, offs3, border, impl=ops.IMPL_GENERIC, relu=relu)
    r = worst(a, g)
    if not r <= 1.0:
        bad += 1
        print(f"BAD fwd2d case {case}: B{B} {H}x{W} shifts {shifts} border {border} relu {relu}: {r:.3f}", flush=True)
    # ---- batch GT, 2-D
    dt = [np.uint8, np.int16, np.int32, np.int64][rng.randint(4)]
    K = int(rng.randint(1, 13))
    offs = [[int(rng.randint(-30, 31)), int(rng.randint(-30, 31))] for _ in range(K)]
    labs = np.stack([O.voronoi_labels((H, W), int(rng.randint(2, 12)), 1000 * case + b, background_frac=0.3, dtype=np.uint8)
                     for b in range(B)]).astype(dt)
    if rng.rand() < 0.2:
        labs[0] = 0
    ignore, padding = bool(rng.rand() < 0.5), bool(rng.rand() < 0.7)
    lab_d = torch.tensor(labs).cuda()
    res = {}
    for route in (0, 1):
        knob(3, route)
        res[route] = ops.gt_targets(lab_d[:, None], ops.offsets3(offs, 2), 0, ignore, padding)
    knob(3, 0)
    same = all(torch.equal(x, y) for x, y in zip(res[0], res[1]))
    to, mo = O.gen_affs_ours(labs[0], offs, ignore=ignore, padding=padding)
    wo = np.stack([O.weight_binary_ratio(c) for c in to])
    ok = same and np.array_equal(res[0][0][0, :, 0].cpu().numpy(), to) and np.array_equal(res[0][1][0, :, 0].cpu().numpy(), mo) \
        and np.array_equal(res[0][2][0, :, 0].cpu().numpy(), wo)
    if not ok:
        bad += 1
        print(f"BAD gt2d case {case}: B{B} {H}x{W} {dt.__name__} K{K} ignore {ignore} padding {padding} routes-equal {same}", flush=True)
    # ---- batch GT, 3-D (seg_to_aff semantics), random nhood
    D = int(rng.randint(1, 9))
    Hs, Ws = int(rng.randint(8, 48)), 4 * int(rng.randint(2, 16))
    nh = np.asarray([[int(rng.randint(-3, 4)), int(rng.randint(-9, 10)), int(rng.randint(-9, 10))] for _ in range(int(rng.randint(1, 13)))], np.int32)
    seg = np.stack([O.voronoi_labels((D, Hs, Ws), 6, 77 * case + b, z_scale=3.0, dtype=np.uint16) for b in range(2)])
    seg = (seg % 250).astype([np.uint8, np.int32, np.uint16][rng.randint(3)])
    seg_d = torch.tensor(seg.astype(np.int32) if seg.dtype == np.uint16 else seg).cuda()
    res = {}
    for route in (0, 1):
        knob(3, route)
        res[route] = ops.gt_targets(seg_d, nh, 1)
    knob(3, 0)
    same = torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][2], res[1][2])
    to = O.seg_to_aff(seg[0], nh, pad='')
    ok = same and np.array_equal(res[0][0][0].cpu().numpy(), to) and \
        np.array_equal(res[0][2][0].cpu().numpy(), np.stack([O.weight_binary_ratio(c) for c in to]))
    if not ok:
        bad += 1
        print(f"BAD gt3d case {case}: D{D} {Hs}x{Ws} {seg.dtype} nhood {nh.tolist()} routes-equal {same}", flush=True)
    # ---- work-list kernel as fallback (fused loss + upstream backward), random 3-D table
    Kf = int(rng.randint(1, 13))
    offf = np.asarray([[int(rng.randint(-2, 3)), int(rng.randint(-6, 7)), int(rng.randint(-6, 7))] for _ in range(Kf)], np.int32)
    e3 = torch.randn((2, 16, D + 2, Hs, Ws), generator=torch.Generator().manual_seed(case + 5)).cuda()
    t3 = (torch.rand((2, Kf, D + 2, Hs, Ws), generator=torch.Generator().manual_seed(case + 6)) > 0.5).float().cuda()
    w3 = (torch.rand((2, Kf, D + 2, Hs, Ws), generator=torch.Generator().manual_seed(case + 7)) + 0.1).cuda()
    norm, coef = np.full(Kf, 1000.0, np.float32), np.linspace(0.5, 1.5, Kf).astype(np.float32)
    bd = ops.BORDER_WRAP if rng.rand() < 0.5 else ops.BORDER_ZERO
    o1 = ops.affs_loss_fwd_bwd(e3, None, offf, bd, t3, w3, None, norm, coef, impl=ops.IMPL_AUTO)
    o2 = ops.affs_loss_fwd_bwd(e3, None, offf, bd, t3, w3, None, norm, coef, impl=ops.IMPL_GENERIC)
    r = [worst(o1[i], o2[i]) for i in range(3)]
    knob(0, 0)
    g1, _ = ops.affs_bwd(e3, None, offf, bd, t3)
    knob(0, 1)
    g2, _ = ops.affs_bwd(e3, None, offf, bd, t3)
    knob(0, 0)
    r.append(worst(g1, g2))
    if not all(x <= 1.0 for x in r):
        bad += 1
        print(f"BAD worklist case {case}: {tuple(e3.shape)} offsets {offf.tolist()} border {bd}: {r}", flush=True)
print(f"stress_r2b: {n_cases} cases, {'FAILED %d' % bad if bad else 'all ok'}", flush=True)
sys.exit(1 if bad else 0)
