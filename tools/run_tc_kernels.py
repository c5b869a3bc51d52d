#!/usr/bin/env python
"""Run the round-2 tcgen05 kernels a few times (for ncu captures): the row-streaming fused inner-product -> regression
kernel at BASELINE config 4's largest point (8 images of (128, 270, 480), D = 192; soft only and soft + arg-extrema) and
at config 2 (32, 64, 144, 240), D = 48, and the tcgen05 adjoint at config 2.  One launch of each per loop."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import realtime_stereo_matcher_b200 as rsm
from realtime_stereo_matcher_b200 import _lib as L

g = torch.Generator(device="cuda").manual_seed(1234)
l4 = (torch.randn((8, 128, 270, 480), device="cuda", generator=g) * 0.5).bfloat16()
r4 = (torch.randn((8, 128, 270, 480), device="cuda", generator=g) * 0.5).bfloat16()
l2 = (torch.randn((32, 64, 144, 240), device="cuda", generator=g) * 0.5).bfloat16()
r2 = (torch.randn((32, 64, 144, 240), device="cuda", generator=g) * 0.5).bfloat16()
go = torch.randn((32, 48, 144, 240), device="cuda", generator=g).bfloat16()
gl, gr = torch.empty_like(l2), torch.empty_like(r2)
lib = L.load()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    rsm.inner_product_regress(l4, r4, 192, argmin=False, argmax=False)
    rsm.inner_product_regress(l4, r4, 192)
    rsm.inner_product_regress(l2, r2, 48, mean=True, argmin=False, argmax=False)
    L.check(lib.rsm_inner_bwd(go.data_ptr(), L.feat(l2), L.feat(r2), gl.data_ptr(), gr.data_ptr(), 32, 64, 144, 240, 48,
                              L.RSM_REDUCE_MEAN, L.dtype_code(l2), L.dtype_code(go), 0, L.stream_ptr(0)), "rsm_inner_bwd")
torch.cuda.synchronize()
print("ok")
