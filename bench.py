#!/usr/bin/env python
"""Benchmark of the cost-volume + disparity-regression hot path (BASELINE.json metric:
stereo pairs/s at 384x1248; cost-volume GB/s vs HBM peak).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (torch port)

One "step" = one pass of the hot path over one batch of synthetic stereo pairs per GPU.  The
default workload is BASELINE config 3 as SURVEY.md F3/8d reads it (the configuration the metric
is quoted on, 8 pairs of 384x1248 per GPU): group-wise correlation (G=8) and concatenate volumes
at MobileStereoNetV4's feature shape (C=32, 1/4 res, D=48) plus the v4 regression head
(trilinear x4 -> softmax over D=192 -> expectation) at full resolution.  Pairs shard across ranks
with no data-path collective (weak scaling).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# --------------------------------------------------------------------------- workloads
class Cfg3:
    """v4 feature shape, 8 pairs/GPU: groupwise(G=8,D=48) + concat(D=48) + v4 head to 384x1248."""
    name = "cfg3: 8 pairs/GPU @384x1248 -- groupwise(G=8)+concat volumes (C=32,1/4 res,D=48) + v4 soft-argmin head (D=192)"
    N, C, H4, W4, D4, G, D, H, W = 8, 32, 96, 312, 48, 8, 192, 384, 1248
    pairs_per_step = 8
    dtype = "f32"
    kernels = ("corr_fwd_kernel[groupwise]", "concat_fwd_kernel", "upsample_regress_fwd_kernel")
    dominant = 1  # index into kernels: concat is the HBM-write stream

    def host_inputs(self, seed, n=None):
        n = n or self.N
        g = torch.Generator().manual_seed(seed)
        return (torch.randn((n, self.C, self.H4, self.W4), generator=g),
                torch.randn((n, self.C, self.H4, self.W4), generator=g),
                torch.randn((n, self.D4, self.H4, self.W4), generator=g) * 3.0)

    def algorithmic_bytes(self):
        """every input element read once, every output element written once."""
        e, n = 4, self.N
        feat = n * self.C * self.H4 * self.W4 * e
        return {
            "groupwise": 2 * feat + n * self.G * self.H4 * self.W4 * self.D4 * e,
            "concat": 2 * feat + 2 * n * self.C * self.H4 * self.W4 * self.D4 * e,
            "v4_head": n * self.D4 * self.H4 * self.W4 * e + n * self.H * self.W * e,
        }

    def step(self, rsm, inp, mark=None):
        left, right, cost = inp
        if mark: mark()
        gw = rsm.groupwise_volume(left, right, self.G, self.D4)
        if mark: mark()
        cat = rsm.concat_volume(left, right, self.D4)
        if mark: mark()
        disp = rsm.v4_head(cost, self.D, self.H, self.W)
        if mark: mark()
        return gw, cat, disp

    def cpu_step(self, tp, inp):
        left, right, cost = inp
        tp.groupwise_volume(left, right, self.G, self.D4)
        tp.concat_volume(left, right, self.D4)
        return tp.v4_tail(cost, self.D, self.H, self.W)


WORKLOADS = {"cfg3": Cfg3}


# ------------------------------------------------------------------------------ helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.count(",") >= 7]
        os.unlink(self.path)
        sm = []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            r = [c.strip() for c in r]
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
            except ValueError:
                continue
            for nme, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if sm:
            out["sm_mhz"] = statistics.median(sm)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full summary."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(kernel)
    return None


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            # host side of the end-to-end path: keep each rank (and the pinned buffers it allocates) on the NUMA
            # node of its own GPU
            from realtime_stereo_matcher_b200.sharding import bind_host_to_device
            bind_host_to_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))   # rank -> GPU stated, not guessed
        else:
            dist.init_process_group(backend)
    return world, rank, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def max_over_ranks(x, world, device):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def time_cpu_port(wl, sample_pairs, budget_s=15.0, max_reps=5):
    """The reference's CPU path (torch port) on a bounded sample; returns (pairs/s, info)."""
    from oracle import torch_port as tp
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    inp = wl.host_inputs(1234, n=sample_pairs)
    with torch.no_grad():
        wl.cpu_step(tp, inp)  # warm-up
        best, spent, reps = float("inf"), 0.0, 0
        while reps < max_reps and (reps == 0 or spent < budget_s):
            t0 = time.perf_counter()
            wl.cpu_step(tp, inp)
            dt = time.perf_counter() - t0
            best, spent, reps = min(best, dt), spent + dt, reps + 1
    return sample_pairs / best, {"cores": threads, "reps": reps, "best_s": best}


# -------------------------------------------------------------------------- reference arm
def run_reference(args, wl):
    world, rank, _ = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), 0
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    from oracle import torch_port as tp
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sample = 1
    inp = wl.host_inputs(1234, n=sample)
    with torch.no_grad():
        for _ in range(args.warmup):
            wl.cpu_step(tp, inp)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            wl.cpu_step(tp, inp)
        dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = f"{sample} of {wl.pairs_per_step} pairs per step (same shapes), torch {torch.__version__} CPU, fp32"
    print(json.dumps({
        "impl": "reference", "metric": "stereo_pairs_per_sec", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
        "config": {"workload": wl.name, "sample": desc},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=_RESULT_OUT, flush=True)


_RESULT_OUT = sys.stdout   # main() swaps in a private copy of the original stdout


# ------------------------------------------------------------------------------ our arm
def run_b200(args, wl):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    import realtime_stereo_matcher_b200 as rsm
    rsm.load_library()
    world, rank, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    # rotating input sets, resident in HBM before the timed region
    nsets = 4
    sets = [tuple(t.to(dev) for t in wl.host_inputs(1234 + 17 * rank + s)) for s in range(nsets)]
    K, Wm = args.steps, max(args.warmup, 3)
    with torch.no_grad():
        for i in range(Wm):
            out = wl.step(rsm, sets[i % nsets])
        del out
        torch.cuda.synchronize()

        # ---- device-resident timed region: K steps, per-op events for the roofline
        marks = [[torch.cuda.Event(enable_timing=True) for _ in range(len(wl.kernels) + 1)] for _ in range(K)]
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        barrier(world)
        torch.cuda.synchronize()
        t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_beg.record()
        for i in range(K):
            it = iter(marks[i])
            out = wl.step(rsm, sets[i % nsets], mark=lambda it=it: next(it).record())
            del out
        t_end.record()
        torch.cuda.synchronize()
        barrier(world)
        ms_total = max_over_ranks(t_beg.elapsed_time(t_end), world, dev)
        op_ms = [statistics.mean(marks[i][j].elapsed_time(marks[i][j + 1]) for i in range(K))
                 for j in range(len(wl.kernels))]

        # ---- end to end through the public API with HOST buffers (pinned), copies inside the region
        host_sets = [tuple(t.pin_memory() for t in wl.host_inputs(99 + 17 * rank + s)) for s in range(2)]
        host_out = torch.empty((wl.N, wl.H, wl.W), dtype=torch.float32).pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in host_sets[0])
        d2h = host_out.numel() * host_out.element_size()

        def e2e_serial(i):
            inp = tuple(t.to(dev, non_blocking=True) for t in host_sets[i % 2])
            disp = wl.step(rsm, inp)[2]
            host_out.copy_(disp, non_blocking=True)
            torch.cuda.synchronize()      # the caller owns the disparity on the host
            return disp

        for i in range(3):
            e2e_serial(i)
        barrier(world)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            e2e_serial(i)
        torch.cuda.synchronize()
        e2e_serial_s = max_over_ranks(time.perf_counter() - t0, world, dev)

        # the public streaming API: H2D of batch i+1 overlaps the kernels of batch i (double buffer)
        pipe = rsm.HostPipeline(lambda l, r, c: wl.step(rsm, (l, r, c))[2], device=dev, depth=2)
        outs = [host_out, torch.empty_like(host_out).pin_memory()]
        for _ in pipe.run((host_sets[i % 2] for i in range(3)), outs):
            pass
        barrier(world)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_done = sum(1 for _ in pipe.run((host_sets[i % 2] for i in range(K)), outs))
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0, world, dev)
        assert n_done == K
        barrier(world)
        clocks = sampler.stop() if rank == 0 else None   # sampled across the device-timed and e2e regions

    pairs = wl.pairs_per_step * world * K
    value = pairs / (ms_total * 1e-3)
    e2e_value = pairs / e2e_s
    if rank != 0:
        return
    # ---- CPU baseline beside it (rank 0, N=1 only): the reference's CPU path on a bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, info = time_cpu_port(wl, sample_pairs=1)
        cpu = {"value": v, "unit": "pairs/s", "cores": info["cores"], "kind": "port",
               "sample": f"1 of {wl.pairs_per_step} pairs (same shapes), best of {info['reps']} passes "
                         f"({info['best_s']:.2f} s each), torch CPU port of the reference op sequence"}
    ab = wl.algorithmic_bytes()
    names = list(ab)
    dom = wl.dominant
    peak, peak_src = measured_peak()
    achieved = ab[names[dom]] / (op_ms[dom] * 1e-3) / 1e9
    line = {
        "metric": "stereo_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": K,
        "warmup": Wm, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
        "config": {"workload": wl.name, "pairs_per_gpu_per_step": wl.pairs_per_step,
                   "l2": "4 rotating input sets; each step writes 3.3 GB of volumes (>> 126 MB L2)",
                   "parallelism": f"dp{world} (batch sharded, no data-path collective)"},
        "roofline": {"bound": "hbm", "kernel": wl.kernels[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(wl.kernels[dom]), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ab[names[dom]], "avg_launch_ms": op_ms[dom]},
        "kernels": {n: {"ms": op_ms[j], "algorithmic_GBps": ab[n] / (op_ms[j] * 1e-3) / 1e9}
                    for j, n in enumerate(names)},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "serial_value": pairs / e2e_serial_s,
                "note": "public API HostPipeline: per step pinned host features+cost -> H2D -> 3 kernels -> D2H of "
                        "the disparity map, H2D of step i+1 overlapped with the kernels of step i (serial_value: no "
                        "overlap); the volumes stay in HBM for the aggregation network, as in the model"},
        "gpu_launches": len(wl.kernels) * K,
        "clocks": clocks,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: everything written to fd 1 from here on (NCCL's version
    # banner under NCCL_DEBUG, library chatter) is sent to stderr, the result goes to a private copy of stdout
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    wl = WORKLOADS[args.workload]()
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
