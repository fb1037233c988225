"""Device-side counterpart of the patch blending in ``scripts_ac3ac4/data/provider_valid.py`` (the step right after the
hot path in full-volume inference).

The reference copies every patch prediction to the host (22 MB for 12 x 18 x 160 x 160), multiplies it by a Gaussian
window and accumulates into NumPy arrays (``add_vol``, :306-321), then divides by the accumulated window
(``get_results``, :323-340) and hands ``1 - affs`` to the Mutex Watershed (inference.py:196).  ``VolumeBlender`` keeps
the accumulators on the GPU and takes the affinity patch straight from ``inf_embedding_loss_norm5(..., relu=True,
replicate_border=True)`` -- one launch per patch, no device -> host traffic until the final volume is read.
Arithmetic is NumPy's (product rounded, then added; IEEE division): results are bit-identical.
"""
import numpy as np
import torch

from ._lib import lib, check
from ._ops import _ptr, _stream


def get_weight(out_size, sigma=0.2, mu=0.0):
    """The reference's window, computed exactly as there (provider_valid.py:294-304): (1, d, h, w) float32 NumPy."""
    zz, yy, xx = np.meshgrid(np.linspace(-1, 1, out_size[0], dtype=np.float32),
                             np.linspace(-1, 1, out_size[1], dtype=np.float32),
                             np.linspace(-1, 1, out_size[2], dtype=np.float32), indexing='ij')
    dd = np.sqrt(zz * zz + yy * yy + xx * xx)
    weight = 1e-6 + np.exp(-((dd - mu) ** 2 / (2.0 * sigma ** 2)))
    return weight[np.newaxis, ...]


class VolumeBlender:
    """out_affs / weight_map of ``Provider_valid`` on the device.

    channels: number of affinity channels; volume_shape (D, H, W): the (padded) raw volume; out_size (d, h, w): the
    patch extent.  ``add_vol(patch, pos)`` with patch (K, d, h, w) [or (1, K, d, h, w)] CUDA float32 and pos the patch
    origin along the three volume axes; ``get_results(valid_padding)`` returns the blended (K, D', H', W') volume."""

    def __init__(self, channels, volume_shape, out_size, device, sigma=0.2, mu=0.0):
        self.K, self.shape, self.out_size = int(channels), tuple(int(v) for v in volume_shape), tuple(int(v) for v in out_size)
        self.weight_vol = torch.from_numpy(np.ascontiguousarray(get_weight(self.out_size, sigma, mu)[0], dtype=np.float32)).to(device)
        self.reset_output()

    def reset_output(self):
        dev = self.weight_vol.device
        self.out_affs = torch.zeros((self.K,) + self.shape, dtype=torch.float32, device=dev)
        self.weight_map = torch.zeros((1,) + self.shape, dtype=torch.float32, device=dev)

    def add_vol(self, affs_vol, pos):
        if affs_vol.dim() == 5:
            if affs_vol.shape[0] != 1:
                raise ValueError("add_vol takes one patch at a time (the reference squeezes a batch of 1)")
            affs_vol = affs_vol[0]
        if not affs_vol.is_cuda or affs_vol.dtype != torch.float32:
            raise TypeError("affs_vol must be a float32 CUDA tensor: pea_b200 has no CPU path")
        if tuple(affs_vol.shape) != (self.K,) + self.out_size:
            raise ValueError(f"patch shape {tuple(affs_vol.shape)} != {(self.K,) + self.out_size}")
        affs_vol = affs_vol.contiguous()
        d, h, w = self.out_size
        D, H, W = self.shape
        with torch.cuda.device(affs_vol.device):
            rc = lib.pea_blend_add(_ptr(affs_vol), _ptr(self.weight_vol), self.K, d, h, w, _ptr(self.out_affs), _ptr(self.weight_map),
                                   D, H, W, int(pos[0]), int(pos[1]), int(pos[2]), _stream())
        check("pea_blend_add", rc)

    def get_results(self, valid_padding=(0, 0, 0), one_minus=False):
        """out_affs / weight_map cropped by valid_padding on both sides (:323-340); ``one_minus`` returns 1 - that, the
        Mutex Watershed's input (inference.py:196).  Voxels no patch covered divide 0 / 0 -> NaN, as in the reference."""
        pz, py, px = (int(v) for v in valid_padding)
        D, H, W = self.shape
        out = torch.empty((self.K, D - 2 * pz, H - 2 * py, W - 2 * px), dtype=torch.float32, device=self.out_affs.device)
        with torch.cuda.device(out.device):
            rc = lib.pea_blend_finish(_ptr(self.out_affs), _ptr(self.weight_map), self.K, D, H, W, pz, py, px, int(bool(one_minus)),
                                      _ptr(out), _stream())
        check("pea_blend_finish", rc)
        return out
