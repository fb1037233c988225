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
torch.cuda.synchronize()
print("sanitize_small ok", float(l.detach()), float(lx.detach()))
