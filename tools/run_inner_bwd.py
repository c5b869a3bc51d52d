#!/usr/bin/env python
"""Run the cfg2 inner-product backward a few times (for ncu launch lists): bf16 and fp32, left and right gradients."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import realtime_stereo_matcher_b200 as rsm

g = torch.Generator(device="cuda").manual_seed(1234)
for dt in (torch.bfloat16, torch.float32):
    L = torch.randn((32, 64, 144, 240), device="cuda", generator=g).to(dt).requires_grad_(True)
    R = torch.randn((32, 64, 144, 240), device="cuda", generator=g).to(dt).requires_grad_(True)
    out = rsm.inner_product_volume(L, R, 48, mean=True)
    go = torch.randn_like(out)
    for _ in range(3):
        torch.autograd.grad(out, (L, R), go, retain_graph=True)
torch.cuda.synchronize()
print("ok")
