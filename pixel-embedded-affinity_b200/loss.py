"""Criterion classes with the reference's names (scripts_*/loss/loss.py:106-152).

``WeightedMSE`` is the criterion every shipped config trains with (cvppp.yaml:22, ac3ac4.yaml loss_func);
passing an instance of it (or of the reference's own class of the same name) to the loss functions of this
package selects the fused CUDA kernel, which reproduces its arithmetic -- including the
``prod(pred.size()[2:]) * pred.size()[0]`` normaliser -- inside the kernel.  Any other criterion is applied
by PyTorch on top of the CUDA affinity op, offset by offset, exactly like the reference loop.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class WeightedMSE(nn.Module):
    """sum(weight * (pred - target)^2) / (prod(pred.shape[2:]) * pred.shape[0])   (loss.py:113-120)."""

    def forward(self, pred, target, weight=None):
        norm = torch.prod(torch.tensor(pred.size()[2:]).float()) * pred.size()[0]
        norm = norm.to(pred.device)
        sq = (pred - target) ** 2
        return torch.sum(sq if weight is None else weight * sq) / norm


class MSELoss(nn.Module):
    """loss.py:126-132 -- the weight argument is ignored."""

    def __init__(self):
        super().__init__()
        self.criterion = nn.MSELoss()

    def forward(self, pred, target, weight=None):
        return self.criterion(pred, target)


class BCELoss(nn.Module):
    """loss.py:134-140"""

    def __init__(self):
        super().__init__()
        self.criterion = nn.BCELoss()

    def forward(self, pred, target, weight=None):
        return self.criterion(pred, target)


class WeightedBCE(nn.Module):
    """loss.py:142-152"""

    def __init__(self, size_average=True, reduce=True):
        super().__init__()
        self.size_average = size_average
        self.reduce = reduce

    def forward(self, pred, target, weight=None):
        return F.binary_cross_entropy(pred, target, weight)


def is_weighted_mse(criterion):
    """True for this package's WeightedMSE and for the reference's class of the same name."""
    return type(criterion).__name__ == "WeightedMSE"


def is_plain_mse(criterion):
    """True for the MSELoss wrapper (this package's or the reference's, loss.py:126-132): ``nn.MSELoss()`` with the mean
    reduction, weight ignored -- i.e. the fused weighted-MSE kernels with unit weights and the element count as normaliser."""
    inner = getattr(criterion, "criterion", None)
    return type(criterion).__name__ == "MSELoss" and isinstance(inner, nn.MSELoss) and inner.reduction == "mean"
