"""Drop-in for ``weight_binary_ratio`` of ``scripts_*/data/data_segmentation.py:205-228``.

    from pea_b200.data_segmentation import weight_binary_ratio        # scripts_cvppp/data/data_provider.py:15
"""
import numpy as np
import torch

from . import _ops as ops
from ._host import device


def weight_binary_ratio(label, mask=None, alpha=1.0):
    """Class-balance weights of one channel: all-equal input -> ones; else with wf = clip(nonzero fraction,
    0.05, 0.99) in float64: wf > 0.5 -> 1 on non-zeros and alpha*wf/(1-wf) on zeros, otherwise
    alpha*(1-wf)/wf on non-zeros and 1 on zeros; float32 out.  NumPy in -> NumPy out; CUDA tensor in -> tensor."""
    if mask is not None:
        raise NotImplementedError("weight_binary_ratio(mask=...) is never used on the reference's hot path "
                                  "(data_provider.py:216-225 pass mask=None) and is not implemented on the device")
    was_numpy = not isinstance(label, torch.Tensor)
    if was_numpy:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(label), dtype=np.float32)).to(device())
    else:
        if not label.is_cuda:
            raise ValueError("label tensors must be CUDA tensors (pass a NumPy array for host data)")
        t = label.float().contiguous()
    w = ops.weight_binary_ratio(t, 1, alpha) if t.numel() else torch.empty_like(t)
    return w.cpu().numpy() if was_numpy else w
