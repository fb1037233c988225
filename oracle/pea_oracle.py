"""CPU oracle for the Pixel-Embedded-Affinity hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement of the reference algorithm (weih527/Pixel-Embedded-Affinity)
for the one path this repo accelerates.  It is imported ONLY by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs,
and only as the checker / the CPU arm -- never by the product package
(``pixel-embedded-affinity_b200``), which has no CPU path at all.

Parity pinning: the reference ships no golden vectors or tests, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build
container by ``tests/golden/make_golden.py`` (which imports the real reference from
``/root/reference``) and committed as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks every function below against those fixtures.

The loss functions are written with torch CPU ops (the reference's own arithmetic is
``F.normalize`` / ``torch.roll`` / ``torch.sum`` and autograd), so the oracle's backward
is autograd through the same op sequence; ``closed_form_grad_2d/3d`` restate the
hand-derived backward in float64 NumPy as an independent check.
The ground-truth generators are integer/byte work and are written in NumPy.

Reference citations (file:line under /root/reference):
  2D loss      scripts_cvppp/loss/loss_embedding_mse.py:7-95
  3D loss      scripts_ac3ac4/loss/loss_embedding_mse.py:7-67,143-289
  criterion    scripts_cvppp/loss/loss.py:106-124        (WeightedMSE)
  offsets/GT   scripts_cvppp/utils/affinity_ours.py:4-39
  3D GT        scripts_cvppp/data/data_affinity.py:22-102
  weights      scripts_cvppp/data/data_segmentation.py:205-228
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

SHIFTS_NORM5 = [1, 1, 1, 2, 3, 3, 3, 9, 9, 4, 27, 27]  # scripts_ac3ac4/loss/loss_embedding_mse.py:176


# ----------------------------------------------------------------------------------------
# criterion  (scripts_cvppp/loss/loss.py:106-124)
# ----------------------------------------------------------------------------------------
class WeightedMSE(torch.nn.Module):
    """sum(w * (pred - target)^2) / (prod(pred.shape[2:]) * pred.shape[0]).

    The reference moves the normaliser with ``.cuda()``; here it simply follows ``pred``.
    The normaliser is a float32 product of the trailing sizes times the batch size, which
    is why a 3-D ``(B, H, W)`` prediction is normalised by ``B * W``.
    """

    def forward(self, pred, target, weight=None):
        trailing = torch.prod(torch.tensor(pred.size()[2:]).float())
        norm_term = (trailing * pred.size()[0]).to(pred.device)
        sq = (pred - target) ** 2
        if weight is not None:
            sq = weight * sq
        return torch.sum(sq) / norm_term


# ----------------------------------------------------------------------------------------
# 2D: embedding -> affinities -> loss   (scripts_cvppp/loss/loss_embedding_mse.py)
# ----------------------------------------------------------------------------------------
def _pair_affinity_2d(anchor, partner, offset, mode):
    """a[b,y,x] = <anchor[b,:,y,x], partner[b,:,(y+dy)%H,(x+dx)%W]>   (:7-13, :49-56, :68-74)."""
    moved = torch.roll(partner, shifts=(-int(offset[0]), -int(offset[1])), dims=(2, 3))
    if mode == 'ours':
        return torch.sum(moved * anchor, dim=1)
    return torch.nn.CosineSimilarity(dim=1, eps=1e-6)(moved, anchor)


def embedding2affs(embedding, offsets, mode='ours'):
    """:58-66 -- normalise, then one circularly-shifted dot product per offset."""
    if mode == 'ours':
        embedding = F.normalize(embedding, p=2, dim=1)
    B, _, H, W = embedding.shape
    affs = torch.zeros((B, len(offsets), H, W), dtype=embedding.dtype, device=embedding.device)
    for k, off in enumerate(offsets):
        affs[:, k] = _pair_affinity_2d(embedding, embedding, off, mode)
    return affs


def embedding_loss(embedding, target, weightmap, mask, criterion, offsets, affs0_weight=1, mode='ours'):
    """:18-47 -- returns (loss, affs with grad, all_loss list of floats).

    ``affs0_weight`` is accepted and ignored, exactly as in the reference (:26-33 build a
    factor table that :38 never applies)."""
    if mode == 'ours':
        embedding = F.normalize(embedding, p=2, dim=1)
    mask = mask.float()
    affs = torch.zeros_like(target)
    loss = torch.tensor(0, dtype=embedding.dtype, device=embedding.device)
    all_loss = []
    for k, off in enumerate(offsets):
        a = _pair_affinity_2d(embedding, embedding, off, mode)
        lk = criterion(a * mask[:, k], target[:, k] * mask[:, k], weightmap[:, k])
        loss = loss + lk
        all_loss.append(lk.item())
        affs[:, k] = a
    return loss, affs, all_loss


def ema_embedding_loss(embedding, ema_embedding, target, weightmap, mask, criterion, offsets,
                       affs0_weight=1, mode='ours'):
    """:79-95 -- two-view variant; offsets 0 and 1 are scaled by ``affs0_weight``."""
    if mode == 'ours':
        embedding = F.normalize(embedding, p=2, dim=1)
        ema_embedding = F.normalize(ema_embedding, p=2, dim=1)
    mask = mask.float()
    affs = torch.zeros_like(target)
    loss = torch.tensor(0, dtype=embedding.dtype, device=embedding.device)
    for k, off in enumerate(offsets):
        a = _pair_affinity_2d(embedding, ema_embedding, off, mode)
        lk = criterion(a * mask[:, k], target[:, k] * mask[:, k], weightmap[:, k])
        loss = loss + (lk * affs0_weight if k < 2 else lk)
        affs[:, k] = a
    return loss, affs


# ----------------------------------------------------------------------------------------
# 3D: axis-aligned affinities   (scripts_ac3ac4/loss/loss_embedding_mse.py)
# ----------------------------------------------------------------------------------------
def _crop(t, axis, lo, hi):
    idx = [slice(None)] * t.dim()
    idx[2 + axis] = slice(lo, hi)
    return t[tuple(idx)]


def _pair_affinity_3d(anchor, partner, axis, shift):
    """a[p] = <anchor[p], partner[p - shift*e_axis]> on the valid region only (:143-151, :237-245)."""
    L = anchor.shape[2 + axis]
    return torch.sum(_crop(anchor, axis, shift, None) * _crop(partner, axis, 0, L - shift),
                     dim=1, keepdim=True)


def _loss3d(embedding, partner, target, weightmap, criterion, shifts, affs0_weight, weighted_upto):
    affs = torch.zeros_like(target)
    loss = 0
    for i, s in enumerate(shifts):
        axis = i % 3
        a = _pair_affinity_3d(embedding, partner, axis, s)
        li = criterion(a, _crop(target[:, i:i + 1], axis, s, None), _crop(weightmap[:, i:i + 1], axis, s, None))
        loss = loss + (li * affs0_weight if i < weighted_upto else li)
        _crop(affs[:, i:i + 1], axis, s, None).copy_(a.detach())
    return loss, affs


def embedding_loss_norm5(embedding, target, weightmap, criterion, affs0_weight=1, shift=1, fill=True):
    """:169-194 -- 12 channels, axis = i % 3, shifts from SHIFTS_NORM5, first three weighted."""
    e = F.normalize(embedding, p=2, dim=1)
    return _loss3d(e, e, target, weightmap, criterion, SHIFTS_NORM5, affs0_weight, 3)


def embedding_loss_norm1(embedding, target, weightmap, criterion, affs0_weight=1, shift=1, fill=True):
    """:7-27 -- 3 channels, one shift, only the z channel weighted by ``affs0_weight``."""
    e = F.normalize(embedding, p=2, dim=1)
    return _loss3d(e, e, target, weightmap, criterion, [shift] * 3, affs0_weight, 1)


def ema_embedding_loss_norm5(embedding, ema_embedding, target, weightmap, criterion, affs0_weight=1,
                             shift=1, fill=True):
    """:263-289"""
    e = F.normalize(embedding, p=2, dim=1)
    u = F.normalize(ema_embedding, p=2, dim=1)
    return _loss3d(e, u, target, weightmap, criterion, SHIFTS_NORM5, affs0_weight, 3)


def ema_embedding_loss_norm1(embedding, ema_embedding, target, weightmap, criterion, affs0_weight=1,
                             shift=1, fill=True):
    """:30-51"""
    e = F.normalize(embedding, p=2, dim=1)
    u = F.normalize(ema_embedding, p=2, dim=1)
    return _loss3d(e, u, target, weightmap, criterion, [shift] * 3, affs0_weight, 1)


def _inf3d(embedding, shifts):
    e = F.normalize(embedding, p=2, dim=1)
    B, _, D, H, W = e.shape
    affs = torch.zeros((B, len(shifts), D, H, W), dtype=e.dtype, device=e.device)
    for i, s in enumerate(shifts):
        axis = i % 3
        _crop(affs[:, i:i + 1], axis, s, None).copy_(_pair_affinity_3d(e, e, axis, s).detach())
    return affs


def inf_embedding_loss_norm5(embedding):
    """:212-234"""
    return _inf3d(embedding, SHIFTS_NORM5)


def inf_embedding_loss_norm1(embedding, shift=1):
    """:54-67"""
    return _inf3d(embedding, [shift] * 3)


# ----------------------------------------------------------------------------------------
# offsets + 2D ground truth   (scripts_cvppp/utils/affinity_ours.py)
# ----------------------------------------------------------------------------------------
def gen_offsets(shift, neighbor=4):
    """:4-9"""
    assert neighbor == 4 or neighbor == 8, 'neigbor must be 4 or 8!'
    base = [[-shift, 0], [0, -shift]]
    if neighbor == 8:
        base += [[-shift, -shift], [-shift, shift]]
    return base


def multi_offset(shifts, neighbor=4):
    """:11-15"""
    out = []
    for s in shifts:
        out += gen_offsets(s, neighbor=neighbor)
    return out


def _neighbour_2d(labels, off):
    """labels[y+dy, x+dx] where in bounds, 0 elsewhere, plus the in-bounds indicator.

    Restates ``scipy.ndimage.shift(labels, -off, order=0, prefilter=False)`` with the default
    ``mode='constant', cval=0`` (:22-25)."""
    H, W = labels.shape
    dy, dx = int(off[0]), int(off[1])
    moved = np.zeros_like(labels)
    inb = np.zeros(labels.shape, dtype=bool)
    y0, y1 = max(0, -dy), min(H, H - dy)
    x0, x1 = max(0, -dx), min(W, W - dx)
    if y0 < y1 and x0 < x1:
        moved[y0:y1, x0:x1] = labels[y0 + dy:y1 + dy, x0 + dx:x1 + dx]
        inb[y0:y1, x0:x1] = True
    return moved, inb


def gen_affs_ours(labels, offsets=[[-1, 0], [0, -1]], ignore=False, padding=False):
    """:17-39 -- (affinities float32 (K,H,W), masks uint8 (K,H,W))."""
    K = len(offsets)
    affinities = np.zeros((K,) + labels.shape, dtype=np.float32)
    masks = np.zeros((K,) + labels.shape, dtype=np.uint8)
    for k, off in enumerate(offsets):
        moved, inb = _neighbour_2d(labels, off)
        same = (labels == moved)            # out-of-bounds neighbour compares against 0
        if ignore:
            same = same & (labels != 0) & (moved != 0)
        same = np.where(inb, same, bool(padding))
        affinities[k] = same.astype(np.float32)
        masks[k] = inb.astype(np.uint8)
    return affinities, masks


# ----------------------------------------------------------------------------------------
# 3D / 2D ground truth by slicing   (scripts_cvppp/data/data_affinity.py)
# ----------------------------------------------------------------------------------------
def mknhood3d(radius=1):
    """:22-40 for radius 1 evaluates to the three negative unit steps (SURVEY.md a11 [probed])."""
    if radius != 1:
        raise NotImplementedError("oracle restates mknhood3d only for radius=1 (the only use on the path)")
    return np.ascontiguousarray(np.array([[-1, 0, 0], [0, -1, 0], [0, 0, -1]], dtype=np.int32))


NHOOD_NORM5 = np.array([[-1, 0, 0], [0, -1, 0], [0, 0, -1],
                        [-2, 0, 0], [0, -3, 0], [0, 0, -3],
                        [-3, 0, 0], [0, -9, 0], [0, 0, -9],
                        [-4, 0, 0], [0, -27, 0], [0, 0, -27]], dtype=np.int32)


def seg_to_aff(seg, nhood=None, pad='replicate'):
    """:53-102 -- aff[e,p] = (seg[p]==seg[p+n_e]) & seg[p]>0 & seg[p+n_e]>0 where both in bounds, else 0;
    with exactly 3 (3-D) or 2 (2-D) edges and pad=='replicate' the first slice along each axis is
    overwritten by seg>0 (:94-100)."""
    if nhood is None:
        nhood = mknhood3d(1)
    nhood = np.asarray(nhood)
    shape = seg.shape
    nd = len(shape)
    E = nhood.shape[0]
    aff = np.zeros((E,) + shape, dtype=np.float32)
    for e in range(E):
        here, there = [], []
        empty = False
        for a in range(nd):
            n = int(nhood[e, a])
            lo, hi = max(0, -n), min(shape[a], shape[a] - n)
            if lo >= hi:
                empty = True
            here.append(slice(lo, hi))
            there.append(slice(lo + n, hi + n))
        if empty:
            continue
        a_, b_ = seg[tuple(here)], seg[tuple(there)]
        aff[(e,) + tuple(here)] = (a_ == b_) * (a_ > 0) * (b_ > 0)
    if pad == 'replicate' and ((nd == 3 and E == 3) or (nd == 2 and E == 2)):
        aff[0, 0] = (seg[0] > 0).astype(aff.dtype)
        aff[1, :, 0] = (seg[:, 0] > 0).astype(aff.dtype)
        if nd == 3:
            aff[2, :, :, 0] = (seg[:, :, 0] > 0).astype(aff.dtype)
    return aff


def seg_to_aff_norm5(seg):
    """The 12-channel target of scripts_ac3ac4/data/data_provider_labeled_deep.py:246-257."""
    return np.concatenate([seg_to_aff(seg, NHOOD_NORM5[0:3], pad=''), seg_to_aff(seg, NHOOD_NORM5[3:6], pad=''),
                           seg_to_aff(seg, NHOOD_NORM5[6:9], pad=''), seg_to_aff(seg, NHOOD_NORM5[9:12], pad='')],
                          axis=0)


# ----------------------------------------------------------------------------------------
# class-balance weights   (scripts_cvppp/data/data_segmentation.py:205-228)
# ----------------------------------------------------------------------------------------
def weight_binary_ratio(label, mask=None, alpha=1.0):
    if label.max() == label.min():
        return np.ones_like(label, np.float32)
    label = (label != 0).astype(int)
    if mask is None:
        wf = float(label.sum()) / np.prod(label.shape)
    else:
        wf = float((label * mask).sum()) / mask.sum()
    wf = np.clip(wf, a_min=5e-2, a_max=0.99)
    if wf > 0.5:
        weight = label + alpha * wf / (1 - wf) * (1 - label)
    else:
        weight = alpha * (1 - wf) / wf * label + (1 - label)
    if mask is not None:
        weight = weight * mask
    return weight.astype(np.float32)


# ----------------------------------------------------------------------------------------
# closed-form backward in float64 NumPy -- independent of autograd
# ----------------------------------------------------------------------------------------
def _normalise64(x):
    x = x.astype(np.float64)
    n = np.maximum(np.sqrt((x * x).sum(axis=1, keepdims=True)), 1e-12)
    return x / n, n


def closed_form_grad_2d(embedding, target, weightmap, mask, offsets, partner=None, affs0_weight=1.0):
    """d loss / d embedding for embedding_loss (partner None) or ema_embedding_loss (partner given,
    treated as constant).  NumPy float64; returns (loss, affs, grad)."""
    x = np.asarray(embedding)
    e, n = _normalise64(x)
    u = e if partner is None else _normalise64(np.asarray(partner))[0]
    B, C, H, W = x.shape
    t = np.asarray(target, dtype=np.float64)
    w = np.asarray(weightmap, dtype=np.float64)
    m = np.asarray(mask, dtype=np.float64)
    G = np.zeros_like(e)
    affs = np.zeros_like(t)
    loss = 0.0
    N = float(B * W)
    for k, (dy, dx) in enumerate(offsets):
        c = affs0_weight if (partner is not None and k < 2) else 1.0
        un = np.roll(u, shift=(-dy, -dx), axis=(2, 3))            # u[p+o]
        a = (e * un).sum(axis=1)
        affs[:, k] = a
        loss += c * (w[:, k] * (a * m[:, k] - t[:, k] * m[:, k]) ** 2).sum() / N
        g = 2.0 * c * w[:, k] * m[:, k] * m[:, k] * (a - t[:, k]) / N
        G += g[:, None] * un
        if partner is None:
            G += np.roll(g[:, None] * e, shift=(dy, dx), axis=(2, 3))  # pair anchored at p-o
    grad = (G - e * (e * G).sum(axis=1, keepdims=True)) / n
    grad = np.where(np.sqrt((x.astype(np.float64) ** 2).sum(axis=1, keepdims=True)) > 1e-12, grad, G / 1e-12)
    return loss, affs, grad


def closed_form_grad_3d(embedding, target, weightmap, shifts, weighted_upto, partner=None, affs0_weight=1.0):
    """Same for the 3-D family (embedding_loss_norm1/5 and the ema variants)."""
    x = np.asarray(embedding)
    e, n = _normalise64(x)
    u = e if partner is None else _normalise64(np.asarray(partner))[0]
    B = x.shape[0]
    t = np.asarray(target, dtype=np.float64)
    w = np.asarray(weightmap, dtype=np.float64)
    G = np.zeros_like(e)
    affs = np.zeros_like(t)
    loss = 0.0
    for i, s in enumerate(shifts):
        ax = 2 + (i % 3)
        L = x.shape[ax]
        if s >= L:
            continue
        c = affs0_weight if i < weighted_upto else 1.0
        hi = [slice(None)] * 5
        lo = [slice(None)] * 5
        hi[ax] = slice(s, None)
        lo[ax] = slice(0, L - s)
        hi, lo = tuple(hi), tuple(lo)
        a = (e[hi] * u[lo]).sum(axis=1)
        hi_k = (hi[0], i) + hi[2:]
        affs[hi_k] = a
        N = float(B * np.prod(a.shape[1:]))
        d = a - t[hi_k]
        loss += c * (w[hi_k] * d * d).sum() / N
        g = 2.0 * c * w[hi_k] * d / N
        G[hi] += g[:, None] * u[lo]
        if partner is None:
            G[lo] += g[:, None] * e[hi]
    grad = (G - e * (e * G).sum(axis=1, keepdims=True)) / n
    return loss, affs, grad


# ----------------------------------------------------------------------------------------
# seeded synthetic inputs shared by tests / bench
# ----------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------------------------------
# undoing the second view's flip augmentation: scripts_cvppp/data/data_consistency.py:19-45 (2-D, rule of size 3) and
# scripts_ac3ac4/utils/consistency_aug.py:58-77, 217-228 (3-D, rule of size 4)
def simple_augment_reverse(data, rule):
    """data (C,H,W) with rule (x reflection, y reflection, xy transpose), or (C,D,H,W) with rule (z reflection,
    x reflection, y reflection, xy transpose): transpose first, then the reflections, as the reference orders them."""
    if data.dim() == 3:
        assert np.size(rule) == 3
        if rule[2]:
            data = data.permute(0, 2, 1)
        if rule[1]:
            data = torch.flip(data, [1])
        if rule[0]:
            data = torch.flip(data, [2])
        return data
    assert np.size(rule) == 4 and data.dim() == 4
    if rule[3]:
        data = data.permute(0, 1, 3, 2)
    if rule[2]:
        data = torch.flip(data, [2])
    if rule[1]:
        data = torch.flip(data, [3])
    if rule[0]:
        data = torch.flip(data, [1])
    return data


def convert_consistency_flip(gt, rules):
    """data_consistency.py:34-45 / consistency_aug.py:217-228."""
    gt = gt.detach().clone()
    rules = np.asarray(rules.data.cpu().numpy() if isinstance(rules, torch.Tensor) else rules).astype(np.uint8)
    return torch.stack([simple_augment_reverse(gt[k], rules[k]) for k in range(gt.shape[0])], dim=0)


# ------------------------------------------------------------------------------------------------------------------
# full-volume inference epilogue: scripts_ac3ac4/data/provider_valid.py:294-340 (get_weight, add_vol, get_results)
def blend_get_weight(out_size, sigma=0.2, mu=0.0):
    """provider_valid.py:294-304 (the num_z >= 18 branch): (1, d, h, w) float32."""
    zz, yy, xx = np.meshgrid(np.linspace(-1, 1, out_size[0], dtype=np.float32),
                             np.linspace(-1, 1, out_size[1], dtype=np.float32),
                             np.linspace(-1, 1, out_size[2], dtype=np.float32), indexing='ij')
    dd = np.sqrt(zz * zz + yy * yy + xx * xx)
    weight = 1e-6 + np.exp(-((dd - mu) ** 2 / (2.0 * sigma ** 2)))
    return weight[np.newaxis, ...]


def blend_add_vol(out_affs, weight_map, affs_vol, weight_vol, pos):
    """provider_valid.py:306-316 ('superhuman' branch), in place; pos = origin along the three volume axes."""
    a, b, c = pos
    d, h, w = affs_vol.shape[1:]
    out_affs[:, a:a + d, b:b + h, c:c + w] += affs_vol * weight_vol
    weight_map[:, a:a + d, b:b + h, c:c + w] += weight_vol


def blend_get_results(out_affs, weight_map, valid_padding=(0, 0, 0)):
    """provider_valid.py:323-340."""
    out = out_affs / weight_map
    pz, py, px = valid_padding
    if pz:
        out = out[:, pz:-pz]
    if py:
        out = out[:, :, py:-py]
    if px:
        out = out[:, :, :, px:-px]
    return out


def voronoi_labels(shape, n_seeds, seed, background_frac=0.3, z_scale=1.0, dtype=np.uint8):
    """Seeded Voronoi instance map with a background (0) region; ids >= 1."""
    rng = np.random.RandomState(seed)
    nd = len(shape)
    pts = rng.rand(n_seeds, nd) * np.asarray(shape, dtype=np.float64)
    grids = np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape], indexing='ij')
    scale = np.ones(nd, dtype=np.float32)
    if nd == 3:
        scale[0] = z_scale
    best = np.full(shape, np.inf, dtype=np.float32)
    lab = np.zeros(shape, dtype=np.int64)
    for i, p in enumerate(pts):
        d = sum(((g - np.float32(p[a])) * scale[a]) ** 2 for a, g in enumerate(grids))
        take = d < best
        best[take] = d[take]
        lab[take] = i + 1
    n_bg = int(round(background_frac * n_seeds))
    if n_bg > 0:
        bg_ids = rng.choice(np.arange(1, n_seeds + 1), size=n_bg, replace=False)
        lab[np.isin(lab, bg_ids)] = 0
    return lab.astype(dtype)
