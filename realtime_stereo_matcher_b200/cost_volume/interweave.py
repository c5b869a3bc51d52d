import torch.nn as nn

from .. import functional as F_rsm


class TorchInterweaveCost(nn.Module):
    """Mirror of reference cost_volume/interweave.py:5-25."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)

    def forward(self, left, right):
        """(N,C,H,W) x2 -> (N,2C,H,W): even channels = left, odd channels = right."""
        return F_rsm.interweave(left, right)

    def __str__(self):
        return self.__class__.__name__
