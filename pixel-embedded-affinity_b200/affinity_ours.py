"""Drop-in for ``scripts_cvppp/utils/affinity_ours.py`` (identical in scripts_bbbc039v1).

    from pea_b200.affinity_ours import multi_offset, gen_affs_ours        # scripts_cvppp/main.py:35, data_provider.py:14
"""
import numpy as np
import torch

from . import _ops as ops
from ._host import labels_to_device


def gen_offsets(shift, neighbor=4):
    """affinity_ours.py:4-9 -- [[-s,0],[0,-s]] (+ the two diagonals for neighbor=8)."""
    assert neighbor == 4 or neighbor == 8, 'neigbor must be 4 or 8!'
    out = [[-shift, 0], [0, -shift]]
    if neighbor == 8:
        out += [[-shift, -shift], [-shift, shift]]
    return out


def multi_offset(shifts, neighbor=4):
    """affinity_ours.py:11-15 -- config (DATA.shifts, DATA.neighbor) -> runtime offset table."""
    out = []
    for shift in shifts:
        out += gen_offsets(shift, neighbor=neighbor)
    return out


def gen_affs_ours(labels, offsets=[[-1, 0], [0, -1]], ignore=False, padding=False):
    """affinity_ours.py:17-39 -- labels (H,W) -> (affinities float32 (K,H,W), masks uint8 (K,H,W)).

    A leading batch dimension is also accepted: (B,H,W) -> (B,K,H,W).  NumPy in -> NumPy out; CUDA tensor in ->
    CUDA tensors out."""
    t, was_numpy = labels_to_device(labels)
    if t.dim() not in (2, 3):
        raise ValueError(f"labels must be (H,W) or (B,H,W), got shape {tuple(t.shape)}")
    batched = t.dim() == 3
    affs, mask = ops.gt_affs2d(t if batched else t[None], offsets, ignore, padding)
    if not batched:
        affs, mask = affs[0], mask[0]
    if was_numpy:
        return affs.cpu().numpy(), mask.cpu().numpy()
    return affs, mask
