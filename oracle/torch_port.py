"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's CPU code path for the hot path,
restated op for op with torch CPU tensors (the reference's arithmetic *is* ATen, so this is what
`bench.py --impl reference` and the `cpu_baseline` leg time on the GPU box's host cores, with all
the threads torch can use).  Not product code; never imported by realtime_stereo_matcher_b200.

Each function keeps the reference's op sequence -- allocate the filled volume, then one
slice-assign per disparity -- because that sequence (D passes over the output, a materialised
product per disparity) is what its CPU time is made of:
  concat        cost_volume/concatenate.py:27-40      groupwise  cost_volume/groupwise.py:12-55
  inner / corr  cost_volume/inner_product.py:29-41, model/mobile_disp_net_c.py:191-204
  difference    model/mobile_stereo_net.py:13-24      interweave cost_volume/interweave.py:13-21
  regression    model/mobile_stereo_net.py:144-147    v4 head    model/mobile_stereo_net_v4.py:512-518
Checked against the golden vectors in tests/test_torch_port.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _shifted(left, right, d):
    """(left[..., d:], right[..., :W-d]): the overlap of the two rows at disparity d."""
    w = left.shape[-1]
    return left[..., d:], right[..., : w - d]


def concat_volume(left, right, max_disparity):
    n, c, h, w = left.shape
    vol = left.new_zeros((n, 2 * c, h, w, max_disparity))
    for d in range(min(max_disparity, w)):
        a, b = _shifted(left, right, d)
        vol[:, :c, :, d:, d] = a
        vol[:, c:, :, d:, d] = b
    return vol.contiguous()


def interweave(left, right):
    n, c, h, w = left.shape
    out = left.new_zeros((n, 2 * c, h, w))
    out[:, 0::2] = left
    out[:, 1::2] = right
    return out.contiguous()


def inner_product_volume(left, right, max_disparity, mean=False):
    n, c, h, w = left.shape
    vol = left.new_zeros((n, max_disparity, h, w))
    for d in range(min(max_disparity, w)):
        a, b = _shifted(left, right, d)
        prod = a * b
        vol[:, d, :, d:] = prod.mean(dim=1) if mean else prod.sum(dim=1)
    return vol.contiguous()


def groupwise_volume(left, right, n_groups, max_disparity):
    n, c, h, w = left.shape
    assert c % n_groups == 0, f"groupwise cost channel ({c}) % #groups ({n_groups}) != 0."
    vol = torch.zeros((n, n_groups, h, w, max_disparity))
    for d in range(min(max_disparity, w)):
        a, b = _shifted(left, right, d)
        vol[:, :, :, d:, d] = (a * b).view(n, n_groups, c // n_groups, h, w - d).mean(dim=2)
    return vol.contiguous()


def difference_volume(left, right, max_disp):
    n, c, h, w = left.shape
    vol = left.new_ones((n, c, max_disp, h, w))
    for d in range(min(max_disp, w)):
        a, b = _shifted(left, right, d)
        vol[:, :, d, :, d:] = a - b
    return vol


def soft_argmax(cost, keepdim=False):
    p = F.softmax(cost, dim=1)
    d = torch.arange(cost.shape[1], dtype=p.dtype).view(1, -1, 1, 1)
    return torch.sum(p * d, dim=1, keepdim=keepdim)


def hard_argmin(cost):
    return torch.argmin(cost, dim=1)


def v4_tail(cost, maxdisp, out_h, out_w):
    fine = F.interpolate(cost.unsqueeze(1), [maxdisp, out_h, out_w], mode="trilinear").squeeze(1)
    return soft_argmax(fine)
