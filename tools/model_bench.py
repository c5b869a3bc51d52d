#!/usr/bin/env python
"""Full-model throughput on the GPU: the reference's own models (baseline/_ref, built by model.build_model from the
reference's configs, seeded random weights, eval mode) run UNPATCHED (stock PyTorch / cuDNN) and PATCHED
(rsm.patch_reference(fuse=True): cost volume, regression, warps, pre/post steps and -- for v4 -- the whole
per-disparity Conv3d loop on librsm_b200.so) on the same synthetic stereo pairs.  BASELINE configs 1-3 are model
inferences: v1 at (1,3,384,1248), DispNetC at (32,3,540,960), v4 at (8,3,384,1248).

    python tools/model_bench.py [--quick] > profiles/rNN_model_bench.jsonl

Timing: CUDA events on the current stream around `reps` forwards after warm-up, L2 flushed between forwards
(256 MB memset), clocks sampled with nvidia-smi during the run.  One JSON line per (model, precision).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CASES = [
    ("cfg1 MobileStereoNet v1", "stereo_net_config.json", (1, 3, 384, 1248)),
    ("cfg2 MobileDispNetC", "disp_net_c_config.json", (32, 3, 540, 960)),
    ("cfg3 MobileStereoNetV4", "stereo_net_config_v4.json", (8, 3, 384, 1248)),
    ("MobileStereoNetV4 single pair", "stereo_net_config_v4.json", (1, 3, 384, 1248)),
    ("MobileStereoNetV2", "stereo_net_config_v2.json", (8, 3, 384, 1248)),
    ("MobileStereoNetV3", "stereo_net_config_v3.json", (8, 3, 384, 1248)),
]


def time_forward(fn, reps, flush):
    evs = []
    for _ in range(reps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs]
    return statistics.mean(ts), min(ts)


def measure(ref, rsm, cfg_name, shape, autocast, reps=5, warmup=2):
    """-> dict with ms / pairs-per-second of the unpatched and the patched forward and their max abs difference."""
    cfg = ref.config(cfg_name)
    torch.manual_seed(1234)
    net = ref.model.build_model(cfg["model"]).cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(7)
    left = torch.rand(shape, device="cuda", generator=g) * 255.0
    right = torch.roll(left, -6, 3)
    junk = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    flush = lambda: junk.zero_()

    def fwd():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
            return net(left, right)[-1]

    for _ in range(warmup):
        want = fwd()
    ms_ref, best_ref = time_forward(fwd, reps, flush)
    rsm.patch_reference(fuse=True)
    try:
        for _ in range(warmup):
            got = fwd()
        ms_new, best_new = time_forward(fwd, reps, flush)
    finally:
        rsm.unpatch_reference()
    n = shape[0]
    return {"model": cfg["model"]["type"], "config": cfg_name, "input": list(shape), "precision": "autocast_fp16" if autocast else "fp32(tf32 convs)",
            "unpatched_ms": ms_ref, "patched_ms": ms_new, "unpatched_pairs_per_s": n / (ms_ref * 1e-3),
            "patched_pairs_per_s": n / (ms_new * 1e-3), "speedup": ms_ref / ms_new, "best_unpatched_ms": best_ref,
            "best_patched_ms": best_new, "max_abs_diff": float((got.float() - want.float()).abs().max()),
            "disparity_scale": float(want.float().abs().max()), "reps": reps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    from oracle import ref_loader
    import realtime_stereo_matcher_b200 as rsm
    import bench
    ref = ref_loader.load()
    rsm.load_library()
    sampler = bench.ClockSampler(0)
    sampler.start()
    rows = []
    for name, cfg_name, shape in (CASES[:3] if args.quick else CASES):
        for autocast in (False, True):
            r = measure(ref, rsm, cfg_name, shape, autocast)
            r["case"] = name
            rows.append(r)
            print(json.dumps(r), flush=True)
    print(json.dumps({"clocks": sampler.stop()}), flush=True)


if __name__ == "__main__":
    main()
