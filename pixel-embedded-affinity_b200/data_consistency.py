"""Drop-in for ``convert_consistency_flip`` of ``scripts_cvppp/data/data_consistency.py:34-45`` (4-D ``gt``, rules of
size 3: x reflection, y reflection, xy transpose) and ``scripts_ac3ac4/utils/consistency_aug.py:217-228`` (5-D ``gt``,
rules of size 4: z reflection, x reflection, y reflection, xy transpose) -- the step right before the two-view loss
(main.py:274 / :214).

    from pea_b200.data_consistency import convert_consistency_flip

One device pass, the rules are read on the device (the reference copies them to the host: a synchronisation in the
middle of every training step), result detached like the reference's ``gt.detach().clone()``.
"""
import torch

from ._lib import lib, check
from ._ops import _ptr, _stream, check_embedding


def convert_consistency_flip(gt, rules):
    nd = gt.dim()
    if nd not in (4, 5):
        raise ValueError(f"gt must be (B,C,H,W) or (B,C,D,H,W), got shape {tuple(gt.shape)}")
    gt = check_embedding("gt", gt.detach(), nd)
    B, C = gt.shape[:2]
    D = gt.shape[2] if nd == 5 else 1
    H, W = gt.shape[-2:]
    rules = torch.as_tensor(rules)
    nr = 4 if nd == 5 else 3
    if rules.dim() != 2 or tuple(rules.shape) != (B, nr):
        raise ValueError(f"rules must have shape ({B}, {nr}), got {tuple(rules.shape)}")
    # astype(np.uint8) in the reference: truncation towards zero, then "if rule[i]"
    r = rules.to(device=gt.device).to(torch.uint8).to(torch.int32)
    if nd == 4:   # (rx, ry, t) -> (rz = 0, rx, ry, t)
        r = torch.cat([torch.zeros((B, 1), dtype=torch.int32, device=gt.device), r], dim=1)
    r = r.contiguous()
    if H != W:    # a transposed sample would not stack with the others in the reference either; needs one host look
        if bool(r[:, 3].any()):
            raise ValueError("xy transpose needs H == W")
    out = torch.empty_like(gt)
    with torch.cuda.device(gt.device):
        rc = lib.pea_consistency_flip(_ptr(gt), _ptr(r), B, C, D, H, W, _ptr(out), _stream())
    check("pea_consistency_flip", rc)
    return out
