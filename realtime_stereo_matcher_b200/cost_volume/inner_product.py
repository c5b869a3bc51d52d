import torch.nn as nn

from .. import functional as F_rsm


class TorchInnerProductCost(nn.Module):
    """Mirror of reference cost_volume/inner_product.py:5-45."""

    def __init__(self, max_disparity, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.max_disparity = max_disparity

    def forward(self, left, right):
        """(N,C,H,W) x2 -> (N,D,H,W): sum_c L[c,x] R[c,x-d] for x >= d, zeros elsewhere."""
        return F_rsm.inner_product_volume(left, right, self.max_disparity, mean=False)

    def __str__(self):
        return f"{self.__class__.__name__} | aijk,aijh->ajkh"
