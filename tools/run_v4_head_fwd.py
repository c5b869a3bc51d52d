#!/usr/bin/env python
"""Run the MobileStereoNetV4 head forward at cfg3 a few times (for ncu): soft only, and with arg-extrema."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import realtime_stereo_matcher_b200 as rsm

g = torch.Generator(device="cuda").manual_seed(1234)
cost = torch.randn((8, 48, 96, 312), device="cuda", generator=g) * 4
with torch.no_grad():
    for _ in range(3):
        rsm.v4_head(cost, 192, 384, 1248)
    rsm.upsample_regress(cost, 192, 384, 1248, argmin=True, argmax=True)
torch.cuda.synchronize()
print("ok")
