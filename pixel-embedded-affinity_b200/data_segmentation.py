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
        return _masked(label, mask, alpha)
    was_numpy = not isinstance(label, torch.Tensor)
    if was_numpy:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(label), dtype=np.float32)).to(device())
    else:
        if not label.is_cuda:
            raise ValueError("label tensors must be CUDA tensors (pass a NumPy array for host data)")
        t = label.float().contiguous()
    w = ops.weight_binary_ratio(t, 1, alpha) if t.numel() else torch.empty_like(t)
    return w.cpu().numpy() if was_numpy else w


def _masked(label, mask, alpha):
    """The ``mask is not None`` branch (:215-216, :224-225), never taken on the reference's hot path (data_provider.py:216-225
    pass mask=None): float64 torch ops on the device in the reference's operation order -- bit-identical for 0/1-valued
    masks (whose sums are exact), within an ulp of float32 otherwise (NumPy's pairwise float summation order)."""
    was_numpy = not isinstance(label, torch.Tensor)
    dev = device() if was_numpy else label.device
    lab = torch.as_tensor(np.asarray(label) if was_numpy else label).to(dev)
    msk = torch.as_tensor(np.asarray(mask) if not isinstance(mask, torch.Tensor) else mask).to(dev)
    if bool(lab.max() == lab.min()):
        w = torch.ones(lab.shape, dtype=torch.float32, device=dev)
    else:
        lab = (lab != 0).to(torch.float64)
        m64 = msk.to(torch.float64)
        wf = float((lab * m64).sum()) / float(m64.sum())
        wf = float(np.clip(wf, a_min=5e-2, a_max=0.99))
        if wf > 0.5:
            w = lab + (alpha * wf / (1 - wf)) * (1 - lab)
        else:
            w = (alpha * (1 - wf) / wf) * lab + (1 - lab)
        w = (w * m64).to(torch.float32)
    return w.cpu().numpy() if was_numpy else w
