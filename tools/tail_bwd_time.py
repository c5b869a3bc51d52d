"""Times the MobileStereoNetV4 head (forward, forward with arg-extrema, backward) at cfg3 and cfg5 sizes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import realtime_stereo_matcher_b200 as rsm


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    rsm.load_library()
    for (b, dc, hc, wc) in ((8, 48, 96, 312), (1, 48, 270, 480), (2, 48, 61, 77)):
        d, h, w = dc * 4, hc * 4, wc * 4
        for dt in (torch.float32, torch.bfloat16):
            cost = torch.randn(b, dc, hc, wc, device="cuda", dtype=dt).requires_grad_(True)
            out = rsm.v4_head(cost, d, h, w)
            go = torch.randn_like(out)
            ms = timed(lambda: torch.autograd.grad(out, cost, go, retain_graph=True))
            print(f"v4_head_bwd {dt} {b}x{dc}x{hc}x{wc}: {ms * 1e3:.1f} us", flush=True)
            with torch.no_grad():
                c = cost.detach()
                ms = timed(lambda: rsm.v4_head(c, d, h, w))
                ms2 = timed(lambda: rsm.upsample_regress(c, d, h, w, argmin=True, argmax=True))
            print(f"v4_head_fwd {dt} {b}x{dc}x{hc}x{wc}: {ms * 1e3:.1f} us, with argmin+argmax {ms2 * 1e3:.1f} us", flush=True)


if __name__ == "__main__":
    main()
