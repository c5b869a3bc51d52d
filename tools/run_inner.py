#!/usr/bin/env python
"""Run the cfg2 inner-product ops a few times (for ncu captures): bf16 tcgen05 volume + fused regress, fp32 SIMT."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import realtime_stereo_matcher_b200 as rsm
g = torch.Generator(device="cuda").manual_seed(1234)
L = torch.randn((32, 64, 144, 240), device="cuda", generator=g)
R = torch.randn((32, 64, 144, 240), device="cuda", generator=g)
Lb, Rb = L.bfloat16(), R.bfloat16()
for _ in range(3):
    rsm.make_correlation_volume(Lb, Rb, 48)
    rsm.inner_product_regress(Lb, Rb, 48, mean=True)
    rsm.make_correlation_volume(L, R, 48)
torch.cuda.synchronize()
print("ok")
