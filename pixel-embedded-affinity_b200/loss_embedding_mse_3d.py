"""Drop-in for ``scripts_ac3ac4/loss/loss_embedding_mse.py`` (the functions main.py / inference.py import).

Same names, positional order, defaults and return tuples as the reference (file:line below are under that
file).  One-line switch in ``scripts_ac3ac4/main.py:36-37``:

    from pea_b200.loss_embedding_mse_3d import embedding_loss_norm1, embedding_loss_norm5
    from pea_b200.loss_embedding_mse_3d import ema_embedding_loss_norm1, ema_embedding_loss_norm5

Channel i pairs voxel p with p - shift_i along axis i % 3 (0 = z, 1 = y, 2 = x); outside the valid region the
returned affinity is 0 and nothing contributes to the loss; every channel has its own normaliser
B * prod(cropped D,H,W); the returned affinities are detached (:186-190).
"""
import numpy as np
import torch

from . import _ops as ops
from ._autograd import (Affinities, FusedAffinityLoss, FusedAffinityLossLabels, FusedSelfAndCrossLoss,
                        MultiFusedAffinityLoss, OpSpec)
from ._lib import PeaError
from .loss import is_plain_mse, is_weighted_mse

SHIFTS_NORM5 = [1, 1, 1, 2, 3, 3, 3, 9, 9, 4, 27, 27]  # :176


def offsets_from_shifts(shifts):
    """shift list -> (K,3) offsets: channel i looks back by shifts[i] along axis i % 3."""
    offs = np.zeros((len(shifts), 3), dtype=np.int32)
    for i, s in enumerate(shifts):
        offs[i, i % 3] = -int(s)
    return offs


def _norms_3d(shape5, shifts):
    B, _, D, H, W = shape5
    norms = []
    for i, s in enumerate(shifts):
        ext = [D, H, W]
        ext[i % 3] -= int(s)
        if ext[i % 3] <= 0:
            raise ValueError(f"shift {s} of channel {i} is not smaller than the axis length {ext[i % 3] + int(s)}; "
                             "the reference's loss is undefined (0/0) for such a volume")
        # torch.prod(torch.tensor(size()[2:]).float()) * size()[0], float32 (loss.py:114-116)
        norms.append(np.prod(np.asarray(ext, dtype=np.float32), dtype=np.float32) * np.float32(B))
    return np.asarray(norms, dtype=np.float32)


def _loss3d(embedding, ema_embedding, target, weightmap, criterion, shifts, affs0_weight, weighted_upto):
    embedding = ops.check_embedding("embedding", embedding, 5)
    if ema_embedding is not None:
        ema_embedding = ops.check_embedding("ema_embedding", ema_embedding, 5)
        if ema_embedding.shape != embedding.shape:
            raise ValueError("ema_embedding and embedding shapes differ")
    B, C, D, H, W = embedding.shape
    K = len(shifts)
    target = ops.check_like("target", target, (B, K, D, H, W))
    weightmap = ops.check_like("weightmap", weightmap, (B, K, D, H, W))
    offs3 = offsets_from_shifts(shifts)
    norms = _norms_3d(embedding.shape, shifts)
    coef = np.ones(K, np.float32)
    coef[:weighted_upto] = affs0_weight
    if is_weighted_mse(criterion) or is_plain_mse(criterion):
        # MSELoss (TRAIN.loss_func == 'MSELoss'): mean over the cropped (B, 1, D', H', W') prediction == WeightedMSE's normaliser
        # there (prod(size()[2:]) * size()[0]), with unit weights
        spec = OpSpec(offs3, ops.BORDER_ZERO, norms, coef, affs_differentiable=False)
        wm = weightmap if is_weighted_mse(criterion) else torch.ones_like(target)
        loss, affs, _ = FusedAffinityLoss.apply(embedding, ema_embedding, target, wm, None, spec)
        return loss, affs
    # any other criterion: CUDA affinity op + the reference's per-channel cropped criterion calls (:159-163)
    affs = Affinities.apply(embedding, ema_embedding, offs3, ops.BORDER_ZERO, ops.IMPL_AUTO)
    loss = 0
    for i, s in enumerate(shifts):
        idx = [slice(None), slice(i, i + 1), slice(None), slice(None), slice(None)]
        idx[2 + i % 3] = slice(int(s), None)
        idx = tuple(idx)
        li = criterion(affs[idx], target[idx], weightmap[idx])
        loss = loss + (li * affs0_weight if i < weighted_upto else li)
    return loss, affs.detach()


def embedding_loss_norm5(embedding, target, weightmap, criterion, affs0_weight=1, shift=1, fill=True):
    """:169-194 -- 12 channels; ``shift`` and ``fill`` are accepted and unused, as in the reference."""
    return _loss3d(embedding, None, target, weightmap, criterion, SHIFTS_NORM5, affs0_weight, 3)


def embedding_loss_norm1(embedding, target, weightmap, criterion, affs0_weight=1, shift=1, fill=True):
    """:7-27 -- 3 channels at one shift; only the z channel is scaled by ``affs0_weight`` (:20)."""
    return _loss3d(embedding, None, target, weightmap, criterion, [shift] * 3, affs0_weight, 1)


def embedding_loss_norm1_multi(embeddings, targets, weightmaps, criterion, affs0_weight=1, shift=1, fill=True):
    """The deep-supervision calls of one training step in ONE launch: equals

        [embedding_loss_norm1(e, t, w, criterion, affs0_weight, shift) for e, t, w in zip(embeddings, targets, weightmaps)]

    i.e. scripts_ac3ac4/main.py:227-230 (emd1..emd4 at 18 x 10^2 .. 80^2 against down4..down1).  With ``WeightedMSE`` the
    list (at most 8 volumes) runs through one work-list kernel that also finalizes every loss, and backward scales all
    gradients in one launch; anything else falls back to the calls above.  Returns a list of (loss, affs)."""
    n = len(embeddings)
    if not (len(targets) == len(weightmaps) == n):
        raise ValueError("embeddings, targets and weightmaps must have the same length")
    shifts = [shift] * 3
    offs3 = offsets_from_shifts(shifts)
    embs = [ops.check_embedding("embedding", e, 5) for e in embeddings]
    if not is_weighted_mse(criterion) or not ops.multi_eligible([dict(emb=e, offs3=offs3) for e in embs]):
        return [embedding_loss_norm1(e, t, w, criterion, affs0_weight, shift, fill)
                for e, t, w in zip(embeddings, targets, weightmaps)]
    specs, tg, wg = [], [], []
    for e, t, w in zip(embs, targets, weightmaps):
        B, C, D, H, W = e.shape
        tg.append(ops.check_like("target", t, (B, 3, D, H, W)))
        wg.append(ops.check_like("weightmap", w, (B, 3, D, H, W)))
        coef = np.ones(3, np.float32)
        coef[:1] = affs0_weight
        specs.append(OpSpec(offs3, ops.BORDER_ZERO, _norms_3d(e.shape, shifts), coef, affs_differentiable=False))
    flat = MultiFusedAffinityLoss.apply(specs, n, *embs, *([None] * n), *tg, *wg, *([None] * n))
    return [(flat[3 * i], flat[3 * i + 1]) for i in range(n)]


def ema_embedding_loss_norm5(embedding, ema_embedding, target, weightmap, criterion, affs0_weight=1, shift=1,
                             fill=True):
    """:263-289 -- a_i[p] = <e^[p], u^[p - s_i]>."""
    return _loss3d(embedding, ema_embedding, target, weightmap, criterion, SHIFTS_NORM5, affs0_weight, 3)


def ema_embedding_loss_norm1(embedding, ema_embedding, target, weightmap, criterion, affs0_weight=1, shift=1,
                             fill=True):
    """:30-51"""
    return _loss3d(embedding, ema_embedding, target, weightmap, criterion, [shift] * 3, affs0_weight, 1)


def embedding_loss_norm5_self_and_cross(embedding, ema_embedding, target, weightmap, criterion, affs0_weight=1,
                                        coef_self=1.0, coef_cross=1.0):
    """Training-step form of the two full-resolution calls of scripts_ac3ac4/main.py:223-224:

        loss_embedding, pred  = embedding_loss_norm5(embedding, target, weightmap, criterion, affs0_weight)
        loss_embedding_cross, _ = ema_embedding_loss_norm5(embedding, ema_embedding, target, weightmap, criterion, affs0_weight)
        total = coef_self * loss_embedding + coef_cross * loss_embedding_cross

    Returns (total, loss_embedding, pred, loss_embedding_cross); only ``total`` is differentiable (the parts are values).
    Both kernels write ONE gradient buffer -- the two-view kernel adds into what the one-view kernel left
    (PEA_EPI_GRAD_ACCUMULATE) -- so autograd does not accumulate two (B,16,D,H,W) gradients in a third pass; the second
    view is a constant, as in both shipped training loops.  Other criteria / shapes: the two calls above."""
    if not is_weighted_mse(criterion):
        l0, pred = embedding_loss_norm5(embedding, target, weightmap, criterion, affs0_weight)
        l1, _ = ema_embedding_loss_norm5(embedding, ema_embedding, target, weightmap, criterion, affs0_weight)
        return coef_self * l0 + coef_cross * l1, l0.detach(), pred, l1.detach()
    embedding = ops.check_embedding("embedding", embedding, 5)
    ema_embedding = ops.check_embedding("ema_embedding", ema_embedding, 5).detach()
    if ema_embedding.shape != embedding.shape:
        raise ValueError("ema_embedding and embedding shapes differ")
    B, C, D, H, W = embedding.shape
    target = ops.check_like("target", target, (B, 12, D, H, W))
    weightmap = ops.check_like("weightmap", weightmap, (B, 12, D, H, W))
    offs3 = offsets_from_shifts(SHIFTS_NORM5)
    norms = _norms_3d(embedding.shape, SHIFTS_NORM5)
    coef = np.ones(12, np.float32)
    coef[:3] = affs0_weight
    spec_s = OpSpec(offs3, ops.BORDER_ZERO, norms, coef * np.float32(coef_self), affs_differentiable=False)
    spec_c = OpSpec(offs3, ops.BORDER_ZERO, norms, coef * np.float32(coef_cross), affs_differentiable=False)
    total, pred, _, lo_s, lo_c = FusedSelfAndCrossLoss.apply(embedding, ema_embedding, target, weightmap, spec_s, spec_c)
    # loss_out[0] of each kernel carries coef_self / coef_cross times the loss the reference's call returns

    def part(lo, c):
        if c != 0:
            return lo[0] / float(c)   # no host -> device copy: stays CUDA-graph capturable
        return sum(lo[1 + k] * float(coef[k]) for k in range(12))
    return total, part(lo_s, coef_self), pred, part(lo_c, coef_cross)


def _inf3d(embedding, shifts, z_skip=0, out=None, halo=None, relu=False, replicate_border=False):
    embedding = ops.check_embedding("embedding", embedding, 5, make_contiguous=False)
    with torch.no_grad():
        affs = ops.affs_fwd(embedding.detach(), None, offsets_from_shifts(shifts), ops.BORDER_ZERO, z_skip=z_skip, out=out,
                            halo=halo, relu=relu)
        if replicate_border:
            if not affs.is_contiguous():
                raise ValueError("replicate_border needs a contiguous output (not a z-range view)")
            ops.replicate_first(affs, 1)
        return affs


def inf_embedding_loss_norm5(embedding, relu=False, replicate_border=False):
    """:212-234 -- forward only, (B,12,D,H,W).

    The two keyword arguments (default off: the reference's result) fold in what every caller does next
    (scripts_ac3ac4/inference.py:160-165, main.py:232-237): ``replicate_border`` copies slice / row / column 1 of the
    three shift-1 channels over slice / row / column 0 (where the affinity is undefined), ``relu`` clamps at zero --
    the clamp rides on the kernel's store, the border copy is one small launch over three thin slabs, instead of
    two more passes over the whole (B,12,D,H,W) result."""
    return _inf3d(embedding, SHIFTS_NORM5, relu=relu, replicate_border=replicate_border)


def inf_embedding_loss_norm1(embedding, shift=1, relu=False, replicate_border=False):
    """:54-67 -- forward only, (B,3,D,H,W).  ``relu`` / ``replicate_border``: see ``inf_embedding_loss_norm5``."""
    return _inf3d(embedding, [shift] * 3, relu=relu, replicate_border=replicate_border)


def embedding_loss_norm5_labels(embedding, seg, criterion, affs0_weight=1):
    """Training-loop form of ``embedding_loss_norm5``: takes the segmentation (B,D,H,W) instead
    of (target, weightmap) and equals

        target[b] = seg_to_aff(seg[b], nhood = the 12 offsets, pad='')              (data/data_affinity.py:53-102)
        weightmap[b, k] = weight_binary_ratio(target[b, k])                         (data/data_segmentation.py:205-228)
        embedding_loss_norm5(embedding, target, weightmap, criterion, affs0_weight) (:169-194)

    i.e. what ``data_provider_labeled_deep.py:246-285`` prepares on the host per sample.  With ``WeightedMSE`` and
    W % 4 == 0 the two (B,12,D,H,W) GT tensors never exist: one counting pass over the segmentation, then the fused
    kernel derives target and weight from the int32 labels it stages next to the embedding.  Returns (loss, affs)."""
    embedding = ops.check_embedding("embedding", embedding, 5)
    B, C, D, H, W = embedding.shape
    if seg.dim() != 4 or tuple(seg.shape) != (B, D, H, W) or not seg.is_cuda:
        raise ValueError(f"seg must be a CUDA tensor of shape {(B, D, H, W)}, got {tuple(seg.shape)}")
    # int32 staging in the kernel: only label dtypes whose every value fits take the fused route (see the 2-D function)
    if is_weighted_mse(criterion) and W % 4 == 0 and seg.dtype in ops.INT32_SAFE_LABELS:
        lab = seg.contiguous()
        lab32 = lab if lab.dtype == torch.int32 else lab.to(torch.int32)
        offs3 = offsets_from_shifts(SHIFTS_NORM5)
        coef = np.ones(12, np.float32)
        coef[:3] = affs0_weight
        spec = OpSpec(offs3, ops.BORDER_ZERO, _norms_3d(embedding.shape, SHIFTS_NORM5), coef, affs_differentiable=False)
        try:
            wtab = ops.gt_pair_weights(lab32, offs3, 1)
            loss, affs, _ = FusedAffinityLossLabels.apply(embedding, lab32, wtab, False, spec, 1)
            return loss, affs
        except PeaError as ex:
            if ex.code != -8:  # PEA_E_MODE: shape the tiled kernel does not take -- two-step route below
                raise
    from .gt import make_targets_3d
    target, weight = make_targets_3d(seg)
    return embedding_loss_norm5(embedding, target, weight, criterion, affs0_weight)
