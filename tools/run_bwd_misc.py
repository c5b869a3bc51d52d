#!/usr/bin/env python
"""Run the cfg3 backward ops a few times (for ncu): group-wise and concatenate volume gradients, v4 head gradient."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import realtime_stereo_matcher_b200 as rsm

g = torch.Generator(device="cuda").manual_seed(1234)
L = torch.randn((8, 32, 96, 312), device="cuda", generator=g).requires_grad_(True)
R = torch.randn((8, 32, 96, 312), device="cuda", generator=g).requires_grad_(True)
for fn in (lambda: rsm.groupwise_volume(L, R, 8, 48), lambda: rsm.concat_volume(L, R, 48)):
    out = fn()
    go = torch.randn_like(out)
    for _ in range(3):
        torch.autograd.grad(out, (L, R), go, retain_graph=True)
    del out, go
cost = torch.randn((8, 48, 96, 312), device="cuda", generator=g).requires_grad_(True)
out = rsm.v4_head(cost, 192, 384, 1248)
go = torch.randn_like(out)
for _ in range(3):
    torch.autograd.grad(out, cost, go, retain_graph=True)
torch.cuda.synchronize()
print("ok")
