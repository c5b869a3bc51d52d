"""Times the concatenation adjoint at cfg3 / cfg4 sizes (L2 flushed between launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import realtime_stereo_matcher_b200 as rsm


def main():
    rsm.load_library()
    junk = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device="cuda")
    for (n, c, h, w, d) in ((8, 32, 96, 312, 48), (1, 32, 270, 480, 48), (1, 32, 270, 480, 96)):
        for dt in (torch.float32, torch.bfloat16):
            L = torch.randn(n, c, h, w, device="cuda", dtype=dt).requires_grad_(True)
            R = torch.randn(n, c, h, w, device="cuda", dtype=dt).requires_grad_(True)
            vol = rsm.concat_volume(L, R, d)
            go = torch.randn_like(vol)
            fn = lambda: torch.autograd.grad(vol, (L, R), go, retain_graph=True)
            for _ in range(3):
                fn()
            tot = 0.0
            for _ in range(10):
                junk.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            us = tot / 10 * 1e3
            nb = vol.numel() * vol.element_size() + 2 * L.numel() * L.element_size()
            print(f"concat_bwd {dt} {n}x{c}x{h}x{w} D={d}: {us:.1f} us, {nb / us * 1e-3:.0f} GB/s", flush=True)
            del vol, go


if __name__ == "__main__":
    main()
