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
    target, mask = ops.gt_affs2d(labels.contiguous(), offsets, ignore, padding)
    B, K = target.shape[:2]
    weight = ops.weight_binary_ratio(target, B * K)
    return target, mask, weight


def make_targets_3d(seg, nhood=NHOOD_NORM5, replicate=False):
    """seg (B,D,H,W) CUDA integer tensor -> target (B,K,D,H,W) f32, weightmap (B,K,D,H,W) f32."""
    if seg.dim() != 4 or not seg.is_cuda:
        raise ValueError("seg must be a (B,D,H,W) CUDA tensor")
    nhood = np.asarray(nhood)
    target = ops.gt_affs3d(seg.contiguous(), nhood, 3 if (replicate and nhood.shape[0] == 3) else 0)
    B, K = target.shape[:2]
    weight = ops.weight_binary_ratio(target, B * K)
    return target, weight
