import torch.nn as nn

from .. import functional as F_rsm


class TorchConcatenateCost(nn.Module):
    """Mirror of reference cost_volume/concatenate.py:5-41 (no parameters or buffers)."""

    def __init__(self, max_disparity, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.max_disparity = max_disparity

    def forward(self, left, right):
        """(N,C,H,W) x2 -> (N,2C,H,W,max_disparity): left copied, right shifted by d, zeros for x < d."""
        return F_rsm.concat_volume(left, right, self.max_disparity)
