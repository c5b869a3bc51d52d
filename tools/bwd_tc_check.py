"""Inner-product adjoint on tcgen05 (csrc/rsm_corr_bwd_tc.cu) on the GPU box: parity against the oracle on the same
rounded inputs (both gradients, one gradient only, mean / sum, ragged widths, D = 1 .. 64, 16 .. 128 channels, strided
feature views), then timing through the raw C ABI at the cfg2 sizes with the L2 flushed.

    python tools/bwd_tc_check.py [--no-time]
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import oracle
import realtime_stereo_matcher_b200 as rsm
from realtime_stereo_matcher_b200 import _lib as L
from golden_io import round_to

DT = {"bf16": torch.bfloat16, "fp16": torch.float16}
RT = {"bf16": 2.0 ** -6, "fp16": 2.0 ** -8}
bad = 0
shapes = [(1, 16, 2, 128, 16), (1, 64, 3, 240, 48), (2, 64, 5, 240, 48), (1, 32, 4, 312, 48), (1, 16, 2, 72, 19), (1, 128, 2, 480, 64),
          (1, 48, 2, 136, 1), (1, 16, 2, 8, 24), (3, 32, 7, 96, 64), (1, 16, 1, 520, 33), (2, 192, 2, 264, 40)]
for (n, c, h, w, d) in shapes:
    for dn in ("bf16", "fp16"):
        for mean in (False, True):
            rng = np.random.default_rng(5)
            l = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
            r = round_to(rng.standard_normal((n, c, h, w)).astype(np.float32), dn)
            go = round_to(rng.standard_normal((n, d, h, w)).astype(np.float32), dn)
            gl, gr = oracle.inner_product_volume_bwd(go, l, r, mean=mean)
            lt = torch.from_numpy(l).cuda().to(DT[dn]).requires_grad_(True)
            rt = torch.from_numpy(r).cuda().to(DT[dn]).requires_grad_(True)
            rsm.inner_product_volume(lt, rt, d, mean=mean).backward(torch.from_numpy(go).cuda().to(DT[dn]))
            atol = RT[dn] * np.sqrt(d) * 4 / (c if mean else 1)
            el = np.abs(lt.grad.float().cpu().numpy() - gl) - RT[dn] * np.abs(gl)
            er = np.abs(rt.grad.float().cpu().numpy() - gr) - RT[dn] * np.abs(gr)
            # one side only
            lt2 = torch.from_numpy(l).cuda().to(DT[dn]).requires_grad_(True)
            rsm.inner_product_volume(lt2, rt.detach(), d, mean=mean).backward(torch.from_numpy(go).cuda().to(DT[dn]))
            e1 = np.abs(lt2.grad.float().cpu().numpy() - gl) - RT[dn] * np.abs(gl)
            ok = el.max() <= atol and er.max() <= atol and e1.max() <= atol
            bad += not ok
            print((n, c, h, w, d), dn, "mean" if mean else "sum", "excess err L %.2e R %.2e Lonly %.2e (atol %.2e)" % (el.max(), er.max(), e1.max(), atol),
                  "" if ok else "  <-- FAIL", flush=True)
# strided views (channel slice + width crop)
n, c, h, w, d = 2, 32, 3, 248, 48
rng = np.random.default_rng(9)
lf = round_to(rng.standard_normal((n, c + 16, h, w + 8)).astype(np.float32), "bf16")
rf = round_to(rng.standard_normal((n, c + 16, h, w + 8)).astype(np.float32), "bf16")
go = round_to(rng.standard_normal((n, d, h, w)).astype(np.float32), "bf16")
lt = torch.from_numpy(lf).cuda().bfloat16().requires_grad_(True)
rt = torch.from_numpy(rf).cuda().bfloat16().requires_grad_(True)
rsm.inner_product_volume(lt[:, 8:8 + c, :, :w], rt[:, 8:8 + c, :, :w], d).backward(torch.from_numpy(go).cuda().bfloat16())
gl, gr = oracle.inner_product_volume_bwd(go, lf[:, 8:8 + c, :, :w], rf[:, 8:8 + c, :, :w])
e = max(np.abs(lt.grad.float().cpu().numpy()[:, 8:8 + c, :, :w] - gl).max(), np.abs(rt.grad.float().cpu().numpy()[:, 8:8 + c, :, :w] - gr).max())
ok = e <= RT["bf16"] * np.sqrt(d) * 4 + RT["bf16"] * np.abs(gl).max()
bad += not ok
print("strided views max err %.3e" % e, "" if ok else "  <-- FAIL", flush=True)
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
    return float(np.median(ts))


lib = L.load()
for name, n, c, h, w, d in [("cfg2 C=64", 32, 64, 144, 240, 48), ("cfg2 C=16", 32, 16, 144, 240, 48), ("cfg4 C=64 D=48", 1, 64, 270, 480, 48),
                            ("cfg4 C=128 D=48", 1, 128, 270, 480, 48), ("cfg4x8 C=128 D=48", 8, 128, 270, 480, 48)]:
    for dn in ("bf16", "fp16"):
        lt = torch.randn(n, c, h, w, device="cuda").to(DT[dn])
        rt = torch.randn(n, c, h, w, device="cuda").to(DT[dn])
        go = torch.randn(n, d, h, w, device="cuda").to(DT[dn])
        gl, gr = torch.empty_like(lt), torch.empty_like(rt)
        call = lambda: L.check(lib.rsm_inner_bwd(go.data_ptr(), L.feat(lt), L.feat(rt), L.ptr(gl), L.ptr(gr), n, c, h, w, d, L.RSM_REDUCE_MEAN,
                                                 L.dtype_code(lt), L.dtype_code(go), 0, L.stream_ptr(0)), "rsm_inner_bwd")
        t = timeit(call)
        prof = torch.zeros(16, dtype=torch.int64, device="cuda")
        L.check(lib.rsm_inner_bwd_profile(go.data_ptr(), L.feat(lt), L.feat(rt), L.ptr(gl), L.ptr(gr), n, c, h, w, d, L.RSM_REDUCE_MEAN,
                                          L.dtype_code(lt), L.dtype_code(go), 0, L.stream_ptr(0), prof.data_ptr()), "profile")
        torch.cuda.synchronize()
        p = prof.cpu().numpy().astype(np.float64)
        roles = {"issuer_wait_accum": p[0] / max(p[3], 1), "issuer_wait_band": p[1] / max(p[3], 1), "issuer_wait_atoms": p[2] / max(p[3], 1),
                 "issuer_cycles_per_cta": p[3] / 148, "builder_wait_grad": p[4] / max(p[7], 1), "builder_wait_free": p[5] / max(p[7], 1),
                 "builder_building": p[6] / max(p[7], 1), "epilogue_wait": p[8] / max(p[9], 1)}
        alg = (n * d * h * w + 4 * n * c * h * w) * 2
        print(json.dumps({"case": name, "dtype": dn, "us": round(t, 1), "frac_hbm": round(alg / (t * 1e-6) / 6452.5e9, 3),
                          "useful_TFLOPs": round(4 * n * c * h * w * d / (t * 1e-6) / 1e12, 1),
                          "roles": {k: round(v, 3) for k, v in roles.items()}}), flush=True)
sys.exit(1 if bad else 0)
