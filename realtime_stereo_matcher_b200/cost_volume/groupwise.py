import torch.nn as nn

from .. import functional as F_rsm


class TorchGroupwiseCost(nn.Module):
    """Mirror of reference cost_volume/groupwise.py:5-56.

    Deviation (documented): the reference allocates its output without dtype or
    device and therefore always returns a CPU fp32 tensor; this module returns the volume on
    ``left.device`` in ``left.dtype`` (``out_dtype=torch.float32`` reproduces the fp32 dtype)."""

    def __init__(self, n_groups, max_disparity, *args, out_dtype=None, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.n_groups = n_groups
        self.max_disparity = max_disparity
        self.out_dtype = out_dtype

    def groupwise(self, left, right, n_groups):
        """(N,C,H,W) x2 -> (N,G,H,W): per-group channel mean of left*right."""
        return F_rsm.groupwise_pointwise(left, right, n_groups)

    def forward(self, left, right):
        """(N,C,H,W) x2 -> (N,G,H,W,D); AssertionError when C % n_groups != 0."""
        return F_rsm.groupwise_volume(left, right, self.n_groups, self.max_disparity, self.out_dtype)
