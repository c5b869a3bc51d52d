"""fp32 features on the tensor cores (3xTF32): error against the float64 oracle, next to the SIMT fp32 kernel
(this process) and the opt-in tensor-core path (child process with RSM_TC_FP32=1), plus timings."""
import os, subprocess, sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import oracle, realtime_stereo_matcher_b200 as rsm
tag = "3xTF32" if os.environ.get("RSM_TC_FP32") == "1" else "SIMT"
def f64_volume(l, r, d, mean):
    n, c, h, w = l.shape
    out = np.zeros((n, d, h, w))
    for i in range(min(d, w)):
        out[:, i, :, i:] = (l[:, :, :, i:].astype(np.float64) * r[:, :, :, :w - i]).sum(1)
    return out / c if mean else out
for (n, c, h, w, d) in [(1, 8, 2, 128, 16), (1, 16, 3, 240, 48), (2, 64, 5, 240, 48), (1, 24, 4, 312, 48), (1, 128, 2, 480, 192), (1, 16, 2, 68, 19), (1, 40, 3, 67, 33)]:
    rng = np.random.default_rng(1)
    l = rng.standard_normal((n, c, h, w)).astype(np.float32); r = rng.standard_normal((n, c, h, w)).astype(np.float32)
    L = torch.from_numpy(l).cuda(); R = torch.from_numpy(r).cuda()
    out = rsm.inner_product_volume(L, R, d).cpu().numpy()
    ref = f64_volume(l, r, d, False)
    err = np.abs(out - ref).max()
    so, mi, ma = rsm.inner_product_regress(L * 0.5, R * 0.5, d)
    vol = f64_volume(l * 0.5, r * 0.5, d, True).astype(np.float32)
    es = np.abs(so.cpu().numpy() - oracle.soft_argmax(vol)).max()
    mm = (mi.cpu().numpy() != oracle.hard_argmin(vol)).mean()
    print(tag, (n, c, h, w, d), "vol maxerr vs f64 %.3e (tol %.3e) nan %d | fused soft err %.2e argmin mismatch %.4f" % (
        err, 2e-5 * np.sqrt(c) * np.abs(l).max() * np.abs(r).max(), np.isnan(out).sum(), es, mm), flush=True)
rng = np.random.default_rng(7); n, c, h, w, d = 2, 32, 6, 156, 24
l = (rng.integers(-8, 9, (n, c, h, w)) / 8.0).astype(np.float32); r = (rng.integers(-8, 9, (n, c, h, w)) / 8.0).astype(np.float32)
L = torch.from_numpy(l).cuda(); R = torch.from_numpy(r).cuda()
print(tag, "dyadic volume equal", np.array_equal(rsm.inner_product_volume(L, R, d).cpu().numpy(), oracle.inner_product_volume(l, r, d)))
# non-finite inputs propagate like the reference's multiply-accumulate
l2 = l.copy(); l2[0, 3, 1, 40] = np.inf; l2[1, 5, 2, 90] = np.nan
o2 = rsm.inner_product_volume(torch.from_numpy(l2).cuda(), R, d).cpu().numpy(); r2 = oracle.inner_product_volume(l2, r, d)
print(tag, "non-finite positions equal", np.array_equal(np.isfinite(o2), np.isfinite(r2)))
def timeit(f, it=20):
    for _ in range(3): f()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it * 1e3
for name, (n, c, h, w, d) in {"cfg2 C64 D48": (8, 64, 135, 240, 48), "cfg2 C16 D48": (8, 16, 135, 240, 48), "cfg4 C128 D192": (1, 128, 135, 240, 192), "C32 D48 96x312": (8, 32, 96, 312, 48)}.items():
    L = torch.randn(n, c, h, w, device="cuda"); R = torch.randn(n, c, h, w, device="cuda")
    print(tag, name, "volume %.1f us  fused regress %.1f us" % (timeit(lambda: rsm.inner_product_volume(L, R, d, mean=True)), timeit(lambda: rsm.inner_product_regress(L, R, d))), flush=True)
if tag == "SIMT" and "--child" not in sys.argv:
    subprocess.run([sys.executable, __file__, "--child"], env=dict(os.environ, RSM_TC_FP32="1"))
