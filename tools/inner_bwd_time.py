import sys, torch
sys.path.insert(0, '/root/repo')
import realtime_stereo_matcher_b200 as rsm
rsm.load_library()
def timed(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for (n, c, h, w, d) in ((1, 64, 270, 480, 96), (1, 128, 270, 480, 96), (1, 128, 270, 480, 192), (1, 64, 270, 480, 192), (8, 128, 270, 480, 192), (32, 64, 144, 240, 48)):
    for dt in ((torch.float32, torch.bfloat16) if 'f32' in sys.argv else (torch.bfloat16,)):
        L = torch.randn(n, c, h, w, device='cuda', dtype=dt).requires_grad_(True)
        R = torch.randn(n, c, h, w, device='cuda', dtype=dt).requires_grad_(True)
        out = rsm.inner_product_volume(L, R, d)
        go = torch.randn_like(out)
        us = timed(lambda: torch.autograd.grad(out, (L, R), go, retain_graph=True))
        nb = (n * d * h * w + 4 * n * c * h * w) * L.element_size()
        print(f"inner_bwd {str(dt)[6:]} N={n} C={c} D={d}: {us:.1f} us, {nb / us * 1e-3:.0f} GB/s algorithmic", flush=True)
