"""Small-shape pass over every kernel family for compute-sanitizer (memcheck / racecheck): tiled + generic, 2-D + 3-D,
self + cross, forward-only, z-shard views, GT."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pea_b200 as P
from pea_b200 import _ops as ops
from pea_b200.zshard import ZShard, sharded_affinities
from oracle import pea_oracle as O

rng = np.random.RandomState(0)
crit = P.WeightedMSE()
# 3-D
B, D, H, W = 1, 6, 30, 36
emb = rng.randn(B, 16, D, H, W).astype(np.float32); ema = rng.randn(B, 16, D, H, W).astype(np.float32)
seg = O.voronoi_labels((D, H, W), 8, 1, z_scale=3.0, dtype=np.uint16)
tgt, wgt = P.make_targets_3d(torch.tensor(seg.astype(np.int32)[None]).cuda())
for impl in (ops.IMPL_GENERIC, ops.IMPL_TILED):
    ops.set_default_impl(impl)
    e = torch.tensor(emb).cuda().requires_grad_(True)
    l, a = P.embedding_loss_norm5(e, tgt, wgt, crit); l.backward()
    e2 = torch.tensor(emb).cuda().requires_grad_(True)
    lx, ax = P.ema_embedding_loss_norm5(e2, torch.tensor(ema).cuda(), tgt, wgt, crit); lx.backward()
    f = P.inf_embedding_loss_norm5(torch.tensor(emb).cuda())
ops.set_default_impl(ops.IMPL_AUTO)
sh = ZShard(1, 16, D + 6, H, W, 1, 2, "cuda"); sh.buf.normal_(); sh.world = 1
sharded_affinities(sh)
# 2-D
offsets = P.multi_offset([1, 3, 9], 8)
lab = O.voronoi_labels((40, 44), 6, 2)
emb2 = rng.randn(2, 16, 40, 44).astype(np.float32)
t2, m2, w2 = P.make_targets_2d(torch.tensor(np.stack([lab, lab])).cuda(), offsets)
for impl in (ops.IMPL_GENERIC, ops.IMPL_TILED):
    ops.set_default_impl(impl)
    e = torch.tensor(emb2).cuda().requires_grad_(True)
    l, a, al = P.embedding_loss(e, t2, w2, m2, crit, offsets); (l + a.sum() * 1e-3).backward()
    e3 = torch.tensor(emb2).cuda().requires_grad_(True)
    lx, ax = P.ema_embedding_loss(e3, torch.tensor(emb2[::-1].copy()).cuda(), t2, w2, m2, crit, offsets); lx.backward()
ops.set_default_impl(ops.IMPL_AUTO)
P.gen_affs_ours(lab, offsets, padding=True); P.seg_to_aff(seg); P.weight_binary_ratio(t2[0, 0].cpu().numpy())
# round 2: multi-problem launch (3-D and 2-D), label-driven axis kernel (CVPPP table), flip, patch blending, pyramid GT
embs = [torch.tensor(rng.randn(1, 16, 5, h, h).astype(np.float32)).cuda().requires_grad_(True) for h in (6, 12)]
tw = [(torch.rand(1, 3, 5, h, h).cuda().round(), torch.rand(1, 3, 5, h, h).cuda() + 0.1) for h in (6, 12)]
r = P.embedding_loss_norm1_multi(embs, [t for t, _ in tw], [w for _, w in tw], crit)
(r[0][0] + r[1][0]).backward()
offs27 = P.multi_offset([1, 3, 5, 9, 27], 4)
lab2 = torch.tensor(np.stack([O.voronoi_labels((66, 96), 6, 4), O.voronoi_labels((66, 96), 6, 5)])).cuda()
e27 = torch.tensor(rng.randn(2, 16, 66, 96).astype(np.float32)).cuda().requires_grad_(True)
l27, a27, _ = P.embedding_loss_labels(e27, lab2, crit, offs27); l27.backward()
(t27, m27, w27), downs = P.make_targets_pyramid_2d(lab2[:, :64, :], offs27)
em2 = [torch.tensor(rng.randn(2, 16, 64 // f, 96 // f).astype(np.float32)).cuda().requires_grad_(True) for f in (2, 4)]
ks = (8, 6)
r2 = P.embedding_loss_multi(em2, [d[:, :k] for d, k in zip(downs, ks)], [d[:, k:2 * k] for d, k in zip(downs, ks)],
                            [d[:, 2 * k:] for d, k in zip(downs, ks)], crit, [offs27[:k] for k in ks])
(r2[0][0] + r2[1][0]).backward()
P.convert_consistency_flip(torch.tensor(rng.randn(2, 4, 3, 37, 37).astype(np.float32)).cuda(), torch.tensor([[1, 0, 1, 1], [0, 1, 1, 0]]).cuda())
bl = P.VolumeBlender(12, (D, H, W + 4), (D, H, W), "cuda")
bl.add_vol(P.inf_embedding_loss_norm5(torch.tensor(emb).cuda(), relu=True, replicate_border=True), (0, 0, 0))
bl.add_vol(f, (0, 0, 4))
bl.get_results((0, 2, 2), one_minus=True)
torch.cuda.synchronize()
print("sanitize_small ok", float(l.detach()), float(lx.detach()), float(l27.detach()))
# round 2, second session: batch GT in one cooperative launch (8 and 4 voxels per thread, 8-bit and 32-bit labels, many
# slices per block) and the two-launch route behind the test knob
from pea_b200._lib import lib as _lib
for two_pass in (0, 1):
    _lib.pea_debug_configure(3, two_pass)  # PEA_KNOB_GT_TWO_PASS
    for (Bg, Hg, Wg, dt) in ((3, 40, 64, np.uint8), (2, 36, 44, np.uint8), (2, 24, 40, np.int32)):
        lg = np.stack([O.voronoi_labels((Hg, Wg), 6, 7 + b) for b in range(Bg)]).astype(dt)
        P.make_targets_2d(torch.tensor(lg).cuda(), P.multi_offset([1, 3, 5, 9, 11], 4), ignore=bool(two_pass))
    P.make_targets_3d(torch.tensor(np.stack([seg, seg]).astype(np.int32)).cuda())
    P.make_targets_3d(torch.tensor((np.stack([seg, seg]) % 200).astype(np.uint8)).cuda())
_lib.pea_debug_configure(3, 0)
torch.cuda.synchronize()
print("sanitize_small: done")
# ABI 7: staged z-shard halo (copy engine + per-band flags on a side stream, one compute launch that waits per tile row)
from pea_b200.zshard import staged_halo_affinities, fused_halo_affinities, HALO


class _Slab:
    def __init__(self, own, tail):
        self.own, self.prev_tail, self._stage = own, tail, None

    def staging(self):
        if self._stage is None:
            b, c, _, h, w = self.own.shape
            self._stage = torch.zeros((b, c, HALO, h, w), device="cuda")
            self._side = torch.cuda.Stream()
        return self._stage, self._side


zb = torch.randn(1, 16, 5 + HALO, 40, 64, device="cuda")
zs = _Slab(zb[:, :, HALO:], zb[:, :, :HALO])
for _ in range(2):
    got = staged_halo_affinities(zs, row0=8, band_rows=16)
assert torch.equal(got, fused_halo_affinities(zs.own, zs.prev_tail))
torch.cuda.synchronize()
print("sanitize_small: staged halo ok")
