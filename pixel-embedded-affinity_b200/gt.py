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
    if labels.shape[-1] % 4 == 0:   # fused two-pass kernels: four pixels per thread, 128-bit stores
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
