"""Device-side ground-truth for a whole batch: labels -> (target, mask, weightmap) without leaving the GPU.

This is what the reference's ``Dataset.__getitem__`` computes per sample on the host
(scripts_cvppp/data/data_provider.py:204-225, scripts_ac3ac4/data/data_provider_labeled_deep.py:246-285), done
per (sample, channel) exactly as there (``separate_weight``), so the training loop can upload the label map
only (1-2 B/px instead of 90 B/px of target+weight+mask).
"""
import numpy as np
import torch

from . import _ops as ops
from .data_affinity import NHOOD_NORM5


def make_targets_2d(labels, offsets, ignore=False, padding=True):
    """labels (B,H,W) CUDA integer tensor -> target (B,K,H,W) f32, mask (B,K,H,W) u8, weightmap (B,K,H,W) f32."""
    if labels.dim() != 3 or not labels.is_cuda:
        raise ValueError("labels must be a (B,H,W) CUDA tensor")
    if labels.shape[-1] % 4 == 0:   # fused batch kernel (pea_gt_targets): 128-bit stores, W % 4 == 0
        offs3 = ops.offsets3(offsets, 2)
        target, mask, weight = ops.gt_targets(labels.contiguous()[:, None], offs3, 0, ignore, padding)
        return target[:, :, 0], mask[:, :, 0], weight[:, :, 0]
    target, mask = ops.gt_affs2d(labels.contiguous(), offsets, ignore, padding)
    B, K = target.shape[:2]
    return target, mask, ops.weight_binary_ratio(target, B * K)


def make_targets_3d(seg, nhood=NHOOD_NORM5, replicate=False):
    """seg (B,D,H,W) CUDA integer tensor -> target (B,K,D,H,W) f32, weightmap (B,K,D,H,W) f32."""
    if seg.dim() != 4 or not seg.is_cuda:
        raise ValueError("seg must be a (B,D,H,W) CUDA tensor")
    nhood = np.asarray(nhood)
    if (replicate and nhood.shape[0] == 3) or seg.shape[-1] % 4 != 0:
        # replicate-padded 3-edge targets (deep-supervision scales) or odd widths: unfused kernels
        target = ops.gt_affs3d(seg.contiguous(), nhood, 3 if (replicate and nhood.shape[0] == 3) else 0)
        B, K = target.shape[:2]
        return target, ops.weight_binary_ratio(target, B * K)
    target, _, weight = ops.gt_targets(seg.contiguous(), nhood, 1)
    return target, weight


def _strided(labels, f):
    """Nearest-neighbour down-sampling by an integer factor as the reference's data providers do it
    (cv2.resize(..., fx=1/f, fy=1/f, interpolation=cv2.INTER_NEAREST), scripts_cvppp/data/data_provider.py:199-202,
    scripts_ac3ac4/data/data_provider_labeled_deep.py:224-232): for extents divisible by f that is plain strided
    sampling label[..., ::f, ::f] (tests/test_oracle_golden.py checks it against cv2)."""
    H, W = labels.shape[-2:]
    if H % f or W % f:
        raise ValueError(f"label extents {H} x {W} are not divisible by the pyramid factor {f}")
    return labels[..., ::f, ::f].contiguous()


def make_targets_pyramid_2d(labels, offsets, nb_half=2, factors=(2, 4, 8, 16), counts=(4, 3, 2, 1), ignore=False,
                            padding=True):
    """ONE label upload -> the ground truth of all five scales of a CVPPP / BBBC039V1 training step
    (scripts_cvppp/data/data_provider.py:198-239).

    labels (B,H,W) CUDA integer tensor.  Returns ((target, mask, weightmap) at full resolution,
    [down1, down2, down3, down4]) where down_i = cat([lb_affs_i, weightmap_i, affs_mask_i], dim=1) in float32 is what the
    reference's loader hands to main.py:284-287 (scale 1/factors[i], first nb_half * counts[i] offsets)."""
    full = make_targets_2d(labels, offsets, ignore, padding)
    downs = []
    for f, c in zip(factors, counts):
        t, m, w = make_targets_2d(_strided(labels, f), offsets[:nb_half * c], ignore, padding)
        downs.append(torch.cat([t, w, m.float()], dim=1))
    return full, downs


def make_targets_pyramid_3d(seg, factors=(2, 4, 8, 16)):
    """ONE segmentation upload -> the ground truth of all five scales of an AC3/AC4 training step
    (scripts_ac3ac4/data/data_provider_labeled_deep.py:224-285): 12-channel target / weight at full resolution
    (pad=''), and down_i = cat([seg_to_aff(seg[..., ::f, ::f]) (3 edges, pad='replicate'), its weight map], dim=1) as
    main.py:227-230 slices them (``down[:, :3]``, ``down[:, 3:]``).  seg (B,D,H,W) CUDA integer tensor."""
    full = make_targets_3d(seg)
    downs = []
    for f in factors:
        t, w = make_targets_3d(_strided(seg, f), nhood=np.asarray([[-1, 0, 0], [0, -1, 0], [0, 0, -1]]), replicate=True)
        downs.append(torch.cat([t, w], dim=1))
    return full, downs
