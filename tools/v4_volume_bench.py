#!/usr/bin/env python
"""SURVEY 8f-1, first step, measured: MobileStereoNetV4's per-disparity volume loop (48 x interweave -> 3 strided
Conv3d -> 1x1 conv, model/mobile_stereo_net_v4.py:443-458) as the reference runs it (48 iterations on width-cropped
views) against patch.v4_volume_batched (one rsm_shift_interweave_fwd + the same cuDNN convolutions on a batch of
48*B).  A stand-in module with the reference's layer shapes is used (the reference is not on the GPU box)."""
import json
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import realtime_stereo_matcher_b200 as rsm
from realtime_stereo_matcher_b200.patch import v4_volume_batched


class V4Like(nn.Module):
    def __init__(self, volume_size=48):
        super().__init__()
        self.volume_size, self.num_groups = volume_size, 1
        self.conv3d = nn.Sequential(
            nn.Conv3d(1, 16, kernel_size=(8, 3, 3), stride=[8, 1, 1], padding=[0, 1, 1]), nn.BatchNorm3d(16), nn.ReLU(),
            nn.Conv3d(16, 32, kernel_size=(4, 3, 3), stride=[4, 1, 1], padding=[0, 1, 1]), nn.BatchNorm3d(32), nn.ReLU(),
            nn.Conv3d(32, 16, kernel_size=(2, 3, 3), stride=[2, 1, 1], padding=[0, 1, 1]), nn.BatchNorm3d(16), nn.ReLU())
        self.volume11 = nn.Sequential(nn.Conv2d(16, 1, 1, 1, 0, bias=False), nn.BatchNorm2d(1), nn.ReLU(inplace=True))


def timed(f, iters=5):
    for _ in range(2): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    torch.manual_seed(0)
    net = V4Like().cuda().eval()
    for B in (1, 2):
        fl = torch.randn((B, 32, 96, 312), device="cuda")
        fr = torch.randn((B, 32, 96, 312), device="cuda")
        W = fl.shape[3]

        def loop():
            vol = fl.new_zeros([B, net.volume_size, 96, W])
            for i in range(net.volume_size):
                x = rsm.interweave_tensors(fl[:, :, :, i:], fr[:, :, :, : W - i]).unsqueeze(1)
                vol[:, i, :, i:] = net.volume11(net.conv3d(x).squeeze(2)).squeeze(1)
            return vol

        with torch.no_grad():
            a, b = loop(), v4_volume_batched(net, fl, fr)
            err = float((a - b).abs().max())
            t_loop, t_batched = timed(loop), timed(lambda: v4_volume_batched(net, fl, fr))
            t_stack = timed(lambda: rsm.shift_interweave_volume(fl, fr, net.volume_size), 10)
        print(json.dumps({"op": "v4_volume(48 disparities, 96x312 features)", "B": B, "loop_ms": round(t_loop, 3),
                          "batched_ms": round(t_batched, 3), "shift_interweave_ms": round(t_stack, 3),
                          "speedup": round(t_loop / t_batched, 2), "max_abs_diff": err}), flush=True)


if __name__ == "__main__":
    main()
