"""Drop-in for ``scripts_cvppp/loss/loss_embedding_mse.py`` and ``scripts_bbbc039v1/loss/loss_embedding_mse.py``.

Same function names, positional order, defaults and return tuples as the reference (file:line below are under
``scripts_cvppp/loss/loss_embedding_mse.py``); the work is done by the sm_100a kernels behind
``include/pea_b200.h``.  One-line switch in ``scripts_cvppp/main.py:31``:

    from pea_b200.loss_embedding_mse_2d import embedding_loss, ema_embedding_loss
"""
import numpy as np
import torch

from . import _ops as ops
from ._autograd import Affinities, FusedAffinityLoss, FusedAffinityLossLabels, LossList, MultiFusedAffinityLoss, OpSpec
from ._lib import PeaError
from .loss import is_plain_mse, is_weighted_mse


def _check_mode(mode):
    """mode == 'ours' (every shipped config: cvppp.yaml:25, bbbc039v1.yaml dis_mode) takes the CUDA kernels; any other value
    selects the reference's second branch, nn.CosineSimilarity(dim=1, eps=1e-6) on the UN-normalised embeddings
    (:11-13, :53-55, :73-75), composed from torch ops on the device (``_cosine_*`` below: complete, not accelerated)."""
    return mode == 'ours'


def _cosine_affs(anchor, partner, offsets):
    cos = torch.nn.CosineSimilarity(dim=1, eps=1e-6)
    return [cos(torch.roll(partner, shifts=(-int(o[0]), -int(o[1])), dims=(2, 3)), anchor) for o in offsets]


def _cosine_loss(embedding, partner, target, weightmap, mask, criterion, offsets, coef):
    """The reference loop (:35-46 / :87-94) in its cosine branch: no normalisation, masked criterion per offset."""
    embedding = ops.check_embedding("embedding", embedding, 4)
    partner = embedding if partner is None else ops.check_embedding("ema_embedding", partner, 4)
    maskf = mask.float()
    per = _cosine_affs(embedding, partner, offsets)
    affs = torch.stack(per, dim=1)
    loss, each = 0, []
    for k, a in enumerate(per):
        lk = criterion(a * maskf[:, k], target[:, k] * maskf[:, k], weightmap[:, k])
        loss = loss + (lk * coef[k] if coef[k] != 1 else lk)
        each.append(lk.detach())
    return loss, affs, LossList(torch.stack(each))


def _norms_2d(B, W, K):
    # WeightedMSE sees a (B, H, W) prediction: prod(size()[2:]) * size()[0] == W * B, in float32 (loss.py:114-116)
    return np.full(K, np.float32(W) * np.float32(B), dtype=np.float32)


def _norms_mean_2d(B, H, W, K):
    # nn.MSELoss (mean): the (B, H, W) prediction's element count
    return np.full(K, np.float32(B * H * W), dtype=np.float32)


def _prep(embedding, target, weightmap, mask, offsets):
    embedding = ops.check_embedding("embedding", embedding, 4)
    B, C, H, W = embedding.shape
    offs3 = ops.offsets3(offsets, 2)
    K = offs3.shape[0]
    target = ops.check_like("target", target, (B, K, H, W))
    weightmap = ops.check_like("weightmap", weightmap, (B, K, H, W))
    mask = ops.check_like("mask", mask, (B, K, H, W), dtypes=(torch.float32, torch.uint8))
    return embedding, target, weightmap, mask, offs3


def embedding2affs(embedding, offsets, mode='ours', relu=False):
    """:58-66.  (B,C,H,W) -> (B,K,H,W); affs[:,k][y,x] = <e^[y,x], e^[(y+dy)%H,(x+dx)%W]>.
    ``relu=True`` (default off: the reference's result) folds the caller's ``F.relu(pred)`` (scripts_cvppp/main.py:312,
    inference) into the kernel's store (PEA_EPI_RELU); the result is then detached (inference form)."""
    embedding = ops.check_embedding("embedding", embedding, 4)
    if not _check_mode(mode):
        return torch.stack(_cosine_affs(embedding, embedding, offsets), dim=1)
    return Affinities.apply(embedding, None, ops.offsets3(offsets, 2), ops.BORDER_WRAP, ops.IMPL_AUTO,
                            ) if not relu else ops.affs_fwd(embedding.detach(), None, ops.offsets3(offsets, 2), ops.BORDER_WRAP, relu=True)


def _generic_criterion(affs, target, weightmap, mask, criterion, coef):
    """The reference loop (:35-46 / :87-94) for criteria the fused kernel does not implement."""
    maskf = mask.float()
    loss = torch.tensor(0, dtype=affs.dtype, device=affs.device)
    each = []
    for k in range(affs.shape[1]):
        lk = criterion(affs[:, k] * maskf[:, k], target[:, k] * maskf[:, k], weightmap[:, k])
        loss = loss + lk * coef[k] if coef[k] != 1 else loss + lk
        each.append(lk)
    return loss, each


def embedding_loss(embedding, target, weightmap, mask, criterion, offsets, affs0_weight=1, mode='ours'):
    """:18-47.  Returns (loss 0-dim tensor, affs (B,K,H,W) carrying grad, all_loss sequence of K floats).

    ``affs0_weight`` is accepted and ignored, as in the reference (:26-33 vs the commented-out :38)."""
    if not _check_mode(mode):
        e, t, w, m, _ = _prep(embedding, target, weightmap, mask, offsets)
        return _cosine_loss(e, None, t, w, m, criterion, offsets, [1] * len(offsets))
    embedding, target, weightmap, mask, offs3 = _prep(embedding, target, weightmap, mask, offsets)
    B, C, H, W = embedding.shape
    K = offs3.shape[0]
    if is_weighted_mse(criterion):
        spec = OpSpec(offs3, ops.BORDER_WRAP, _norms_2d(B, W, K), np.ones(K, np.float32), affs_differentiable=True)
        loss, affs, each = FusedAffinityLoss.apply(embedding, None, target, weightmap, mask, spec)
        return loss, affs, LossList(each)
    if is_plain_mse(criterion):  # TRAIN.loss_func == 'MSELoss': unit weights, mean over all B * H * W elements
        spec = OpSpec(offs3, ops.BORDER_WRAP, _norms_mean_2d(B, H, W, K), np.ones(K, np.float32), affs_differentiable=True)
        loss, affs, each = FusedAffinityLoss.apply(embedding, None, target, torch.ones_like(target), mask, spec)
        return loss, affs, LossList(each)
    affs = Affinities.apply(embedding, None, offs3, ops.BORDER_WRAP, ops.IMPL_AUTO)
    loss, each = _generic_criterion(affs, target, weightmap, mask, criterion, [1] * K)
    return loss, affs, LossList(torch.stack([e.detach() for e in each]))


def embedding_loss_multi(embeddings, targets, weightmaps, masks, criterion, offsets_list, affs0_weight=1, mode='ours'):
    """The deep-supervision calls of one training step in ONE launch: equals

        [embedding_loss(e, t, w, m, criterion, o) for e, t, w, m, o in zip(embeddings, targets, weightmaps, masks, offsets_list)]

    i.e. scripts_cvppp/main.py:284-287 (emd1..emd4 at 272^2 .. 34^2 with offsets[:8], [:6], [:4], [:2]).  With
    ``WeightedMSE`` and at most 8 scales of at most 12 offsets the list runs through one work-list kernel that also
    finalizes every loss, and backward scales all gradients in one launch; anything else falls back to the calls above.
    Returns a list of (loss, affs, all_loss)."""
    if not _check_mode(mode):
        return [embedding_loss(e, t, w, m, criterion, o, affs0_weight, mode)
                for e, t, w, m, o in zip(embeddings, targets, weightmaps, masks, offsets_list)]
    n = len(embeddings)
    if not (len(targets) == len(weightmaps) == len(masks) == len(offsets_list) == n):
        raise ValueError("embeddings, targets, weightmaps, masks and offsets_list must have the same length")
    prep = [_prep(e, t, w, m, o) for e, t, w, m, o in zip(embeddings, targets, weightmaps, masks, offsets_list)]
    probs = [dict(emb=p[0], offs3=p[4]) for p in prep]
    if not is_weighted_mse(criterion) or not ops.multi_eligible(probs):
        return [embedding_loss(e, t, w, m, criterion, o, affs0_weight, mode)
                for e, t, w, m, o in zip(embeddings, targets, weightmaps, masks, offsets_list)]
    specs = []
    for emb, _, _, _, offs3 in prep:
        B, C, H, W = emb.shape
        K = offs3.shape[0]
        specs.append(OpSpec(offs3, ops.BORDER_WRAP, _norms_2d(B, W, K), np.ones(K, np.float32), affs_differentiable=True))
    flat = MultiFusedAffinityLoss.apply(specs, n, *[p[0] for p in prep], *([None] * n), *[p[1] for p in prep],
                                        *[p[2] for p in prep], *[p[3] for p in prep])
    return [(flat[3 * i], flat[3 * i + 1], LossList(flat[3 * i + 2])) for i in range(n)]


def ema_embedding_loss(embedding, ema_embedding, target, weightmap, mask, criterion, offsets, affs0_weight=1,
                       mode='ours'):
    """:79-95.  Two-view variant: a = <e^[p], u^[p+o]>; offsets 0 and 1 are scaled by ``affs0_weight``.
    Returns (loss, affs)."""
    if not _check_mode(mode):
        e, t, w, m, _ = _prep(embedding, target, weightmap, mask, offsets)
        loss, affs, _ = _cosine_loss(e, ema_embedding, t, w, m, criterion, offsets,
                                     [affs0_weight if k < 2 else 1 for k in range(len(offsets))])
        return loss, affs
    embedding, target, weightmap, mask, offs3 = _prep(embedding, target, weightmap, mask, offsets)
    ema_embedding = ops.check_embedding("ema_embedding", ema_embedding, 4)
    if ema_embedding.shape != embedding.shape:
        raise ValueError(f"ema_embedding shape {tuple(ema_embedding.shape)} != embedding shape {tuple(embedding.shape)}")
    B, C, H, W = embedding.shape
    K = offs3.shape[0]
    coef = np.ones(K, np.float32)
    coef[:2] = affs0_weight
    if is_weighted_mse(criterion):
        spec = OpSpec(offs3, ops.BORDER_WRAP, _norms_2d(B, W, K), coef, affs_differentiable=True)
        loss, affs, _ = FusedAffinityLoss.apply(embedding, ema_embedding, target, weightmap, mask, spec)
        return loss, affs
    if is_plain_mse(criterion):
        spec = OpSpec(offs3, ops.BORDER_WRAP, _norms_mean_2d(B, H, W, K), coef, affs_differentiable=True)
        loss, affs, _ = FusedAffinityLoss.apply(embedding, ema_embedding, target, torch.ones_like(target), mask, spec)
        return loss, affs
    affs = Affinities.apply(embedding, ema_embedding, offs3, ops.BORDER_WRAP, ops.IMPL_AUTO)
    loss, _ = _generic_criterion(affs, target, weightmap, mask, criterion, [affs0_weight if k < 2 else 1 for k in range(K)])
    return loss, affs


def embedding_loss_labels(embedding, labels, criterion, offsets, affs0_weight=1, mode='ours', ignore=False, padding=True,
                          alpha=1.0):
    """Training-loop form of ``embedding_loss``: takes the instance labels (B,H,W) instead of
    (target, weightmap, mask) and equals

        target, mask = gen_affs_ours(labels[b], offsets, ignore, padding)       (utils/affinity_ours.py:17-39)
        weightmap[b, k] = weight_binary_ratio(target[k], alpha=alpha)           (data/data_segmentation.py:205-228)
        embedding_loss(embedding, target, weightmap, mask, criterion, offsets)  (:18-47)

    i.e. what ``data_provider.py:204-225`` prepares on the host per sample.  For the canonical tables
    ``multi_offset(shifts <= 12, neighbor=4)`` with ``WeightedMSE`` the three GT tensors never exist: one counting pass
    over the labels, then the loss kernel derives target / mask / weight itself (172 instead of 258 bytes per pixel,
    and no GT generation pass).  Everything else takes the two-step route.  Returns (loss, affs, all_loss)."""
    embedding = ops.check_embedding("embedding", embedding, 4)
    B, C, H, W = embedding.shape
    if labels.dim() != 3 or tuple(labels.shape) != (B, H, W) or not labels.is_cuda:
        raise ValueError(f"labels must be a CUDA tensor of shape {(B, H, W)}, got {tuple(labels.shape)}")
    offs3 = ops.offsets3(offsets, 2)
    K = offs3.shape[0]
    # the kernel stages int32 labels: only dtypes whose every value fits take the fused route (an int64 / uint32 / uint64
    # id >= 2^31 would wrap and could collide with another id or flip the `> 0` test); the others go two-step, which
    # compares the labels in their own dtype
    if _check_mode(mode) and is_weighted_mse(criterion) and W % 4 == 0 and labels.dtype in ops.INT32_SAFE_LABELS:
        lab = labels.contiguous()
        lab32 = lab if lab.dtype == torch.int32 else lab.to(torch.int32)
        spec = OpSpec(offs3, ops.BORDER_WRAP, _norms_2d(B, W, K), np.ones(K, np.float32), affs_differentiable=True)
        try:
            wtab = ops.gt_pair_weights(lab32[:, None], offs3, 0, ignore, padding, alpha)
            loss, affs, each = FusedAffinityLossLabels.apply(embedding, lab32, wtab, bool(ignore), spec, 0)
            return loss, affs, LossList(each)
        except PeaError as ex:
            if ex.code != -8:  # PEA_E_MODE: not a canonical table / shape -- two-step route below
                raise
    if alpha != 1.0:
        raise NotImplementedError("alpha != 1 is implemented on the fused label-driven path only")
    from .gt import make_targets_2d
    target, mask, weight = make_targets_2d(labels, offsets, ignore, padding)
    return embedding_loss(embedding, target, weight, mask, criterion, offsets, affs0_weight, mode)
