"""Drop-in for ``scripts_*/data/data_affinity.py`` (``seg_to_aff``, ``mknhood3d``; identical in all three trees).

    from pea_b200.data_affinity import seg_to_aff        # scripts_ac3ac4/data/data_provider_labeled_deep.py:29
"""
import numpy as np
import torch

from . import _ops as ops
from ._host import labels_to_device

# The 12-channel neighbourhood of the AC3/AC4 provider (data_provider_labeled_deep.py:246-257), in channel order
NHOOD_NORM5 = np.array([[-1, 0, 0], [0, -1, 0], [0, 0, -1], [-2, 0, 0], [0, -3, 0], [0, 0, -3],
                        [-3, 0, 0], [0, -9, 0], [0, 0, -9], [-4, 0, 0], [0, -27, 0], [0, 0, -27]], dtype=np.int32)


def mknhood3d(radius=1):
    """data_affinity.py:22-40 -- lattice offsets within ``radius``, one per undirected edge."""
    ceilrad = int(np.ceil(radius))
    rng = np.arange(-ceilrad, ceilrad + 1)
    # the reference's meshgrid(z, y, x) ('xy' indexing) ravels with its 2nd input outermost and its 3rd innermost,
    # then stacks columns as (3rd, 1st, 2nd) variable (:33-39)
    cand = [(c, a, b) for b in rng for a in rng for c in rng if a * a + b * b + c * c <= radius * radius]
    half = cand[:len(cand) // 2]
    return np.ascontiguousarray(np.array(half[::-1], dtype=np.int32).reshape(-1, 3))


def seg_to_aff(seg, nhood=None, pad='replicate'):
    """data_affinity.py:53-102 -- seg (D,H,W) or (H,W) -> (E,)+seg.shape float32.

    aff[e,p] = (seg[p] == seg[p+n_e]) & (seg[p] > 0) & (seg[p+n_e] > 0) on in-volume pairs, 0 elsewhere;
    pad == 'replicate' with exactly 3 (3-D) / 2 (2-D) edges overwrites the first slice along axis e of
    channel e with seg > 0.  NumPy in -> NumPy out; CUDA tensor in -> CUDA tensor out."""
    if nhood is None:
        nhood = mknhood3d(1)
    nhood = np.asarray(nhood, dtype=np.int64)
    t, was_numpy = labels_to_device(seg)
    nd = t.dim()
    if nd not in (2, 3):
        raise ValueError(f"seg must be (D,H,W) or (H,W), got shape {tuple(t.shape)}")
    if nhood.ndim != 2 or nhood.shape[1] != nd:
        raise ValueError(f"nhood must have shape (E,{nd}), got {nhood.shape}")
    E = nhood.shape[0]
    rep = nd if (pad == 'replicate' and E == nd) else 0
    if nd == 2:
        nh3 = np.concatenate([np.zeros((E, 1), np.int64), nhood], axis=1)
        aff = ops.gt_affs3d(t[None, None], nh3, rep)[0, :, 0]
    else:
        aff = ops.gt_affs3d(t[None], nhood, rep)[0]
    return aff.cpu().numpy() if was_numpy else aff
