"""Timing of the fused inner-product -> regression op at the cfg2 / cfg4 sizes (bf16), L2 flushed between launches.
Optional argument: NAME=v1,v2 sweeps an environment switch of the library (A/B runs)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, '.')
import realtime_stereo_matcher_b200 as rsm

def timeit(fn, iters=10):
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))

name, vals = (sys.argv[1].split("=")[0], sys.argv[1].split("=")[1].split(",")) if len(sys.argv) > 1 else ("RSM_NONE", ["0"])
cases = [("cfg2 C=64", 32, 64, 144, 240, 48, True), ("cfg2 C=16", 32, 16, 144, 240, 48, True), ("cfg4 C=128 D=192", 1, 128, 270, 480, 192, False),
         ("cfg4 C=64 D=96", 1, 64, 270, 480, 96, False), ("cfg4x8 C=128 D=192", 8, 128, 270, 480, 192, False), ("cfg4x8 C=64 D=96", 8, 64, 270, 480, 96, False)]
for cname, n, c, h, w, d, mean in cases:
    lt = (torch.randn(n, c, h, w, device="cuda") * 0.5).bfloat16()
    rt = (torch.randn(n, c, h, w, device="cuda") * 0.5).bfloat16()
    for am in (False, True):
        row = {"case": cname, "out": "soft+argmin+argmax" if am else "soft"}
        for v in vals:
            os.environ[name] = v
            t = timeit(lambda: rsm.inner_product_regress(lt, rt, d, mean=mean, argmin=am, argmax=am))
            alg = 2 * n * c * h * w * 2 + n * h * w * (4 + (16 if am else 0))
            row[f"{name}={v}"] = {"us": round(t, 1), "frac_hbm": round(alg / (t * 1e-6) / 6452.5e9, 3), "useful_TFLOPs": round(2 * n * c * h * w * d / (t * 1e-6) / 1e12, 1)}
        print(json.dumps(row), flush=True)
