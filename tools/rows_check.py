"""Row-streaming fused inner-product -> regression kernel (csrc/rsm_corr_rows.cu) on the GPU box: parity against the
oracle on small shapes (soft, lse-free; argmin / argmax; dyadic bit-exact; mean / sum; fill region; NaN), then timing
with the per-role cycle counters at the cfg2 / cfg4 sizes next to volume + regression as two launches.

    python tools/rows_check.py [--no-time]
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import oracle
import realtime_stereo_matcher_b200 as rsm
from realtime_stereo_matcher_b200 import _lib as L
from golden_io import round_to

torch.manual_seed(0)
bad = 0


def run(lt, rt, d, mean, argmin=True, argmax=True):
    return rsm.inner_product_regress(lt, rt, d, mean=mean, argmin=argmin, argmax=argmax)


shapes = [(1, 16, 2, 128, 16), (1, 16, 3, 240, 48), (2, 64, 5, 240, 48), (1, 32, 4, 312, 48), (1, 64, 2, 480, 128),
          (1, 16, 2, 72, 19), (1, 128, 2, 480, 192), (1, 32, 3, 200, 130), (2, 16, 2, 304, 260), (1, 48, 2, 136, 1),
          (1, 16, 2, 8, 24), (1, 96, 3, 264, 65), (1, 16, 1, 520, 384), (3, 32, 7, 96, 64)]
for (n, c, h, w, d) in shapes:
    for dn, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        for mean in (False, True):
            rng = np.random.default_rng(1)
            l = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32) * 0.5, dn)
            r = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32) * 0.5, dn)
            lt = torch.from_numpy(l).cuda().to(dt)
            rt = torch.from_numpy(r).cuda().to(dt)
            so, mi, ma = run(lt, rt, d, mean)
            so2, _, _ = run(lt, rt, d, mean, False, False)
            vol = oracle.inner_product_volume(l, r, d, mean=mean)
            es = np.abs(so.cpu().numpy() - oracle.soft_argmax(vol)).max()
            es2 = np.abs(so2.cpu().numpy() - oracle.soft_argmax(vol)).max()
            mm = (mi.cpu().numpy() != oracle.hard_argmin(vol)).mean()
            mx = (ma.cpu().numpy() != oracle.hard_argmax(vol)).mean()
            ok = es <= 1e-4 * d + 1e-5 and es2 <= 1e-4 * d + 1e-5 and mm < 2e-3 and mx < 2e-3
            bad += not ok
            print((n, c, h, w, d), dn, "mean" if mean else "sum", "soft err %.2e / %.2e" % (es, es2),
                  "argmin mismatch %.4f argmax mismatch %.4f" % (mm, mx), "" if ok else "  <-- FAIL", flush=True)

# dyadic inputs: sums are exact, so argmin / argmax must be bit-exact (ties: first index)
for (n, c, h, w, d, mean) in [(2, 32, 6, 160, 24, False), (1, 64, 3, 240, 48, True), (1, 128, 2, 480, 192, False), (1, 16, 2, 320, 100, True)]:
    rng = np.random.default_rng(7)
    l = (rng.integers(-8, 9, (n, c, h, w)) / 8.0).astype(np.float32)
    r = (rng.integers(-8, 9, (n, c, h, w)) / 8.0).astype(np.float32)
    lt = torch.from_numpy(l).cuda().bfloat16()
    rt = torch.from_numpy(r).cuda().bfloat16()
    so, mi, ma = run(lt, rt, d, mean)
    vol = oracle.inner_product_volume(l, r, d, mean=mean)
    e1 = np.array_equal(mi.cpu().numpy(), oracle.hard_argmin(vol))
    e2 = np.array_equal(ma.cpu().numpy(), oracle.hard_argmax(vol))
    bad += not (e1 and e2)
    print("dyadic", (n, c, h, w, d), "mean" if mean else "sum", "argmin equal", e1, "argmax equal", e2, flush=True)

# NaN in the left features inside the x < d fill region must not leak (the reference never computes those entries);
# a NaN inside the band wins both extrema at its first disparity
n, c, h, w, d = 1, 16, 2, 200, 48
rng = np.random.default_rng(3)
l = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), "bf16")
r = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), "bf16")
l[0, 3, 0, 5] = np.nan
r[0, 2, 1, 100] = np.nan
lt = torch.from_numpy(l).cuda().bfloat16()
rt = torch.from_numpy(r).cuda().bfloat16()
so, mi, ma = run(lt, rt, d, False)
vol = oracle.inner_product_volume(l, r, d)
e1 = np.array_equal(mi.cpu().numpy(), oracle.hard_argmin(vol))
e2 = np.array_equal(ma.cpu().numpy(), oracle.hard_argmax(vol))
ref = oracle.soft_argmax(vol)
e3 = np.array_equal(np.isnan(so.cpu().numpy()), np.isnan(ref))
bad += not (e1 and e2 and e3)
print("NaN case: argmin equal", e1, "argmax equal", e2, "NaN positions equal", e3, flush=True)
print("FAILURES:", bad, flush=True)

if "--no-time" in sys.argv:
    sys.exit(1 if bad else 0)


def timeit(fn, iters=10):
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(min(ts))


def profile(lt, rt, d, mean, argmin, argmax):
    n, c, h, w = lt.shape
    so = torch.empty((n, h, w), dtype=torch.float32, device="cuda")
    mi = torch.empty((n, h, w), dtype=torch.int64, device="cuda") if argmin else None
    ma = torch.empty((n, h, w), dtype=torch.int64, device="cuda") if argmax else None
    out = L.RsmRegressOut(L.ptr(so), L.ptr(mi), L.ptr(ma), L.ptr(None), L.ptr(None))
    prof = torch.zeros(16, dtype=torch.int64, device="cuda")
    L.check(L.load().rsm_inner_regress_fwd_profile(L.feat(lt), L.feat(rt), n, c, h, w, d, L.RSM_REDUCE_MEAN if mean else L.RSM_REDUCE_SUM,
                                                  L.dtype_code(lt), out, 0, L.stream_ptr(0), prof.data_ptr()), "profile")
    torch.cuda.synchronize()
    p = prof.cpu().numpy().astype(np.float64)
    return {"issuer_wait_operands": p[0] / max(p[2], 1), "issuer_wait_tmem": p[1] / max(p[2], 1), "issuer_cycles_per_cta": p[2] / 148,
            "producer_wait_slots": p[3] / max(p[4], 1), "epilogue_wait_accumulator": p[5] / max(p[6], 1), "epilogue_scan": p[7] / max(p[6], 1), "epilogue_arrive": p[8] / max(p[6], 1),
            "epilogue_merge": p[9] / max(p[6], 1), "epilogue_cycles_per_warp": p[6] / (148 * 16)}


cases = [("cfg2 C=64", 32, 64, 144, 240, 48, True), ("cfg2 C=16", 32, 16, 144, 240, 48, True),
         ("cfg4 C=128 D=192", 1, 128, 270, 480, 192, False), ("cfg4 C=64 D=96", 1, 64, 270, 480, 96, False),
         ("cfg4 C=32 D=48", 1, 32, 270, 480, 48, False), ("cfg4x8 C=128 D=192", 8, 128, 270, 480, 192, False)]
for name, n, c, h, w, d, mean in cases:
    lt = (torch.randn(n, c, h, w, device="cuda") * 0.5).bfloat16()
    rt = (torch.randn(n, c, h, w, device="cuda") * 0.5).bfloat16()
    row = {"case": name, "shape": [n, c, h, w, d], "dtype": "bf16"}
    for label, am in (("soft", False), ("soft+argmin+argmax", True)):
        med, best = timeit(lambda: run(lt, rt, d, mean, am, am))
        med0, best0 = timeit(lambda: rsm.regress(rsm.inner_product_volume(lt, rt, d, mean=mean), argmin=am, argmax=am))
        alg = 2 * n * c * h * w * 2 + n * h * w * (4 + (16 if am else 0))
        row[label] = {"rows_us": round(med, 1), "rows_best_us": round(best, 1), "volume_then_regress_us": round(med0, 1),
                      "frac_hbm": round(alg / (med * 1e-6) / 6452.5e9, 3), "useful_TFLOPs": round(2 * n * c * h * w * d / (med * 1e-6) / 1e12, 1),
                      "roles": {k: round(v, 3) for k, v in profile(lt, rt, d, mean, am, am).items()}}
    print(json.dumps(row), flush=True)
sys.exit(1 if bad else 0)
