#!/usr/bin/env python
"""Benchmark of the cost-volume + disparity-regression hot path (BASELINE.json metric:
stereo pairs/s at 384x1248; cost-volume GB/s vs HBM peak).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...     # the UNMODIFIED reference's CPU path (baseline/_ref)
    python bench.py --workload cfg5_train --gpus N ...          # training step (BASELINE config 5): fwd + bwd + DDP

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
    l2_note = "4 rotating input sets; each step writes 3.3 GB of volumes (>> 126 MB L2)"

    def host_inputs(self, seed, n=None, dtype=torch.float32):
        n = n or self.N
        g = torch.Generator().manual_seed(seed)
        return tuple(t.to(dtype) for t in (torch.randn((n, self.C, self.H4, self.W4), generator=g),
                                           torch.randn((n, self.C, self.H4, self.W4), generator=g),
                                           torch.randn((n, self.D4, self.H4, self.W4), generator=g) * 3.0))

    def algorithmic_bytes(self, e=4):
        """every input element read once, every output element written once."""
        n = self.N
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

    def ref_step(self, ref, inp):
        """The same step through the UNMODIFIED reference (baseline/_ref) on the CPU: cost_volume/groupwise.py:24-56,
        cost_volume/concatenate.py:11-41, model/mobile_stereo_net_v4.py:511-518."""
        import torch.nn.functional as F
        left, right, cost = inp
        gw = ref.cv_groupwise.TorchGroupwiseCost(self.G, self.D4)(left, right)
        cat = ref.cv_concatenate.TorchConcatenateCost(self.D4)(left, right)
        c = F.interpolate(cost.unsqueeze(1), [self.D, self.H, self.W], mode="trilinear").squeeze(1)
        disp = ref.v4.disparity_regression(F.softmax(c, dim=1), self.D)
        return gw, cat, disp

    def port_step(self, tp, inp):
        left, right, cost = inp
        tp.groupwise_volume(left, right, self.G, self.D4)
        tp.concat_volume(left, right, self.D4)
        return tp.v4_tail(cost, self.D, self.H, self.W)


WORKLOADS = {"cfg3": Cfg3}


def config_dict(wl, world):
    """The SAME dict in both arms (the driver compares it)."""
    return {"workload": wl.name, "pairs_per_gpu_per_step": wl.pairs_per_step, "l2": wl.l2_note,
            "parallelism": f"dp{world} (batch sharded, no data-path collective)"}


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
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "power_w_max": None}
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
        sm, pw = [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            r = [c.strip() for c in r]
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
                pw.append(float(r[2]))
            except ValueError:
                continue
            for nme, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if sm:
            out["sm_mhz"] = statistics.median(sm)
        if pw:
            out["power_w_max"] = max(pw)
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


def load_reference():
    """(kind, callable step(wl, inputs)) for the CPU arm: the unmodified reference from baseline/_ref when it is
    installed (tools/install_ref.py; gpurun ships it), else the torch port of its op sequence."""
    from oracle import ref_loader
    if ref_loader.available():
        ref = ref_loader.load()
        ref_loader.verify_unmodified()
        return "reference", (lambda wl, inp: wl.ref_step(ref, inp))
    from oracle import torch_port as tp
    return "port", (lambda wl, inp: wl.port_step(tp, inp))


def time_cpu_reference(wl, budget_s=20.0, max_reps=3):
    """The reference's CPU path on the full per-GPU batch, best of <= max_reps passes within the budget."""
    kind, step = load_reference()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    inp = wl.host_inputs(1234)
    with torch.no_grad():
        t0 = time.perf_counter()
        step(wl, inp)                                 # warm-up (also the first timing if it alone eats the budget)
        best = time.perf_counter() - t0
        spent, reps = best, 1
        while reps < max_reps and spent < budget_s:
            t0 = time.perf_counter()
            step(wl, inp)
            dt = time.perf_counter() - t0
            best, spent, reps = min(best, dt), spent + dt, reps + 1
    return wl.pairs_per_step / best, {"cores": threads, "reps": reps, "best_s": best, "kind": kind}


# -------------------------------------------------------------------------- reference arm
def run_reference(args, wl):
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    kind, step = load_reference()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    inp = wl.host_inputs(1234)
    with torch.no_grad():
        for _ in range(args.warmup):
            step(wl, inp)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(wl, inp)
        dt = time.perf_counter() - t0
    value = wl.pairs_per_step * args.steps / dt
    what = ("the UNMODIFIED reference (baseline/_ref: TorchGroupwiseCost, TorchConcatenateCost, v4 "
            "F.interpolate->softmax->disparity_regression)" if kind == "reference"
            else "torch CPU port of the reference op sequence (baseline/_ref not installed)")
    desc = (f"all {wl.pairs_per_step} pairs of the per-GPU batch per step, {what}, torch {torch.__version__} CPU, fp32, "
            f"{threads} threads")
    print(json.dumps({
        "impl": "reference", "metric": "stereo_pairs_per_sec", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
        "config": config_dict(wl, max(world, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=_RESULT_OUT, flush=True)


_RESULT_OUT = sys.stdout   # main() swaps in a private copy of the original stdout


# ------------------------------------------------------------------------------ our arm
def time_op(fn, reps, flush):
    """Mean device time of fn() over `reps` launches, L2 flushed before each (CUDA events on the current stream)."""
    evs = []
    for _ in range(reps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return statistics.mean(a.elapsed_time(b) for a, b in evs)


def tensor_core_kernels(rsm, dev, peak):
    """Driver-timed numbers for the tcgen05 path (not part of `value`): BASELINE config 2 -- MobileDispNetC's mean
    correlation at (32, 64, 144, 240), D = 48 in bf16 -- as the volume-writing kernel and as the fused
    build -> regress kernel (no volume)."""
    n, c, h, w, d = 32, 64, 144, 240, 48
    g = torch.Generator(device=dev).manual_seed(5)
    sets = [(torch.randn((n, c, h, w), device=dev, generator=g).to(torch.bfloat16),
             torch.randn((n, c, h, w), device=dev, generator=g).to(torch.bfloat16)) for _ in range(3)]
    junk = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device=dev)     # 512 MB > L2; its memset also covers the host-side launch cost of the next op
    flush = lambda: junk.zero_()
    state = {"i": 0}

    def nxt():
        state["i"] += 1
        return sets[state["i"] % 3]

    out = {}
    feat = 2 * n * c * h * w * 2
    flops = 2.0 * n * c * h * w * d
    # the adjoint through the raw C ABI (both gradients), so that autograd bookkeeping is not in the number
    from realtime_stereo_matcher_b200 import _lib as L
    gout = torch.randn((n, d, h, w), device=dev, generator=g).to(torch.bfloat16)
    gl, gr = torch.empty_like(sets[0][0]), torch.empty_like(sets[0][1])

    def bwd():
        l, r = nxt()
        L.check(L.load().rsm_inner_bwd(gout.data_ptr(), L.feat(l), L.feat(r), gl.data_ptr(), gr.data_ptr(), n, c, h, w, d,
                                       L.RSM_REDUCE_MEAN, L.dtype_code(l), L.dtype_code(gout), dev.index or 0,
                                       L.stream_ptr(dev.index or 0)), "rsm_inner_bwd")

    for name, fn, bytes_, fl in (
            ("inner_mean_fwd[bf16,cfg2,tcgen05]", lambda: rsm.make_correlation_volume(*nxt(), d), feat + n * d * h * w * 2, flops),
            ("inner_regress_fused[bf16,cfg2,tcgen05,soft]",
             lambda: rsm.inner_product_regress(*nxt(), d, mean=True, argmin=False, argmax=False), feat + n * h * w * 4, flops),
            ("inner_regress_fused[bf16,cfg2,tcgen05,soft+argmin+argmax]",
             lambda: rsm.inner_product_regress(*nxt(), d, mean=True), feat + n * h * w * 20, flops),
            ("inner_mean_bwd[bf16,cfg2,tcgen05]", bwd, 2 * feat + n * d * h * w * 2, 2 * flops)):
        for _ in range(3):
            fn()
        ms = time_op(fn, 10, flush)
        out[name] = {"ms": ms, "algorithmic_GBps": bytes_ / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": bytes_ / (ms * 1e-3) / 1e9 / peak,
                     "useful_TFLOPs": fl / (ms * 1e-3) / 1e12}
    # BASELINE config 4's largest point (1, 128, 270, 480), D = 192: the fused kernel where the tensor pipe matters
    n4, c4, h4, w4, d4 = 8, 128, 270, 480, 192
    s4 = [(torch.randn((n4, c4, h4, w4), device=dev, generator=g).to(torch.bfloat16),
           torch.randn((n4, c4, h4, w4), device=dev, generator=g).to(torch.bfloat16)) for _ in range(2)]
    fn = lambda: rsm.inner_product_regress(*s4[state["i"] % 2], d4, argmin=False, argmax=False)
    for _ in range(3):
        fn()
    ms = time_op(fn, 10, flush)
    b4 = 2 * n4 * c4 * h4 * w4 * 2 + n4 * h4 * w4 * 4
    out["inner_regress_fused[bf16,cfg4 C=128 D=192 x8 images,tcgen05,soft]"] = {
        "ms": ms, "algorithmic_GBps": b4 / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": b4 / (ms * 1e-3) / 1e9 / peak,
        "useful_TFLOPs": 2.0 * n4 * c4 * h4 * w4 * d4 / (ms * 1e-3) / 1e12}
    del sets, s4
    out.update(cfg5_regression_numbers(rsm, dev, peak, flush))
    out.update(cfg3_backward_numbers(rsm, dev, peak, flush))
    out.update(v4_model_numbers(rsm, dev, flush))
    return out


def cfg3_backward_numbers(rsm, dev, peak, flush):
    """The training half of BASELINE config 3's step: adjoints of the three kernels of `value` at the same shapes
    (8 pairs, C = 32, 96 x 312, D = 48, G = 8; head to 384 x 1248), fp32, against the HBM roofline (the head's adjoint is
    issue-bound: its figure is fine disparities per second)."""
    n, c, h, w, d, ng = 8, 32, 96, 312, 48, 8
    g = torch.Generator(device=dev).manual_seed(11)
    L = torch.randn((n, c, h, w), device=dev, generator=g).requires_grad_(True)
    R = torch.randn((n, c, h, w), device=dev, generator=g).requires_grad_(True)
    out = {}
    feat = n * c * h * w * 4
    for name, vol_fn, vol_elems in (("groupwise_bwd[f32,cfg3]", lambda: rsm.groupwise_volume(L, R, ng, d), n * ng * h * w * d),
                                    ("concat_bwd[f32,cfg3]", lambda: rsm.concat_volume(L, R, d), n * 2 * c * h * w * d)):
        with torch.enable_grad():
            vol = vol_fn()
        go = torch.randn(vol.shape, device=dev, generator=g)
        fn = lambda: torch.autograd.grad(vol, (L, R), go, retain_graph=True)
        for _ in range(3):
            fn()
        ms = time_op(fn, 10, flush)
        b = vol_elems * 4 + (2 if name.startswith("concat") else 4) * feat     # gradient read, both feature gradients written (+ features read)
        out[name] = {"ms": ms, "algorithmic_GBps": b / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": b / (ms * 1e-3) / 1e9 / peak}
        del vol, go
    cost = (torch.randn((n, d, h, w), device=dev, generator=g) * 3).requires_grad_(True)
    with torch.enable_grad():
        disp = rsm.v4_head(cost, 4 * d, 4 * h, 4 * w)
    go = torch.randn(disp.shape, device=dev, generator=g)
    fn = lambda: torch.autograd.grad(disp, cost, go, retain_graph=True)
    for _ in range(3):
        fn()
    ms = time_op(fn, 10, flush)
    out["v4_head_bwd[f32,cfg3]"] = {"ms": ms, "G_fine_disparities_per_s": n * 4 * d * 4 * h * 4 * w / (ms * 1e-3) / 1e9}
    return out


def v4_model_numbers(rsm, dev, flush):
    """SURVEY 8f-1 / BASELINE config 3 as a MODEL inference: MobileStereoNetV4's per-disparity volume as the fused
    tcgen05 op, and the whole reference model (baseline/_ref, seeded random weights, eval) unpatched vs patched at
    (8,3,384,1248).  Skipped when the reference is not installed."""
    try:
        from oracle import ref_loader
        if not ref_loader.available():
            return {}
        ref = ref_loader.load()
        import importlib.util
        spec = importlib.util.spec_from_file_location("rsm_model_bench", os.path.join(ROOT, "tools", "model_bench.py"))
        mb = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mb)
    except Exception as e:                                   # the bench line must not depend on the checker
        return {"full_model_error": repr(e)[:200]}
    out = {}
    torch.manual_seed(1234)
    net = ref.model.build_model(ref.config("stereo_net_config_v4.json")["model"]).to(dev).eval()
    b = 8
    L, R = torch.randn((b, 32, 96, 312), device=dev), torch.randn((b, 32, 96, 312), device=dev)
    with torch.no_grad():
        fn = lambda: rsm.v4_cost_volume(L, R, net.conv3d, net.volume11, 48)
        for _ in range(2):
            fn()
        ms = time_op(fn, 5, flush)
    px = b * 96 * 312 * 48
    flops = 2.0 * px * (8 * 72 * 16 + 2 * 576 * 32 + 576 * 16 + 16)        # the reference loop's MACs x 2
    out["v4_cost_volume_fused[f32 features, fp16 operands, 8 pairs, tcgen05]"] = {
        "ms": ms, "ms_per_pair": ms / b, "reference_loop_TFLOPs_equivalent": flops / (ms * 1e-3) / 1e12}
    del net, L, R
    for autocast in (False, True):
        r = mb.measure(ref, rsm, "stereo_net_config_v4.json", (8, 3, 384, 1248), autocast, reps=3, warmup=2)
        out[f"full_model_v4_384x1248_b8[{r['precision']}]"] = {
            "unpatched_ms": r["unpatched_ms"], "patched_ms": r["patched_ms"], "unpatched_pairs_per_s": r["unpatched_pairs_per_s"],
            "patched_pairs_per_s": r["patched_pairs_per_s"], "max_abs_diff_px": r["max_abs_diff"], "disparity_scale_px": r["disparity_scale"]}
    # BASELINE configs 1 and 2 as MODEL inferences: MobileStereoNet v1 on one 384x1248 pair, MobileDispNetC
    # on 32 pairs at 540x960 -- the unmodified reference unpatched vs patched, fp32
    for tag, cfg_name, shape in (("cfg1_model_v1_384x1248_b1", "stereo_net_config.json", (1, 3, 384, 1248)),
                                 ("cfg2_model_dispnetc_540x960_b32", "disp_net_c_config.json", (32, 3, 540, 960))):
        r = mb.measure(ref, rsm, cfg_name, shape, False, reps=3, warmup=2)
        out[f"{tag}[{r['precision']}]"] = {
            "unpatched_ms": r["unpatched_ms"], "patched_ms": r["patched_ms"], "unpatched_pairs_per_s": r["unpatched_pairs_per_s"],
            "patched_pairs_per_s": r["patched_pairs_per_s"], "max_abs_diff_px": r["max_abs_diff"], "disparity_scale_px": r["disparity_scale"]}
    return out


def cfg5_regression_numbers(rsm, dev, peak, flush):
    """BASELINE config 5's kernels on one 1080p pair, (1, 192, 1080, 1920) fp32: regression forward (soft + argmin in one
    pass) and backward, against the HBM roofline."""
    n, d, h, w = 1, 192, 1080, 1920
    g = torch.Generator(device=dev).manual_seed(7)
    cost = (torch.randn((n, d, h, w), device=dev, generator=g) * 4).requires_grad_(True)
    out = {}
    fwd = lambda: rsm.regress(cost.detach(), argmin=True, argmax=False)
    for _ in range(3):
        fwd()
    ms = time_op(fwd, 10, flush)
    b = n * d * h * w * 4 + n * h * w * 12
    out["regress_fwd_soft+argmin[f32,cfg5 1080p D=192]"] = {"ms": ms, "algorithmic_GBps": b / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": b / (ms * 1e-3) / 1e9 / peak}
    with torch.enable_grad():
        soft = rsm.soft_argmax(cost)
    gsoft = torch.ones_like(soft)
    bwd = lambda: torch.autograd.grad(soft, cost, gsoft, retain_graph=True)
    for _ in range(3):
        bwd()
    ms = time_op(bwd, 10, flush)
    b = 2 * n * d * h * w * 4 + 3 * n * h * w * 4
    out["soft_argmax_bwd[f32,cfg5 1080p D=192]"] = {"ms": ms, "algorithmic_GBps": b / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": b / (ms * 1e-3) / 1e9 / peak}
    return out


def sustained_concat(rsm, wl, sets, peak, seconds=2.0):
    """The dominant HBM-bound kernel back to back for >= `seconds` of device time: does the roofline fraction hold
    under a long load (power / clocks), not only in a 15 ms burst?"""
    ab = wl.algorithmic_bytes()["concat"]
    reps = max(50, int(seconds / 0.00047))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(reps):
        left, right, _ = sets[i % len(sets)]
        out = rsm.concat_volume(left, right, wl.D4)
        del out
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    gbs = ab / (ms * 1e-3) / 1e9
    return {"kernel": "concat_fwd_kernel", "launches": reps, "seconds": a.elapsed_time(b) * 1e-3, "avg_launch_ms": ms,
            "achieved": gbs, "frac": gbs / peak}


def h2d_ceiling(dev, host_sets, K, world):
    """The box's host->device ceiling for this rank while every rank copies at once: the e2e run's own pinned host
    buffers (same tensors, same alternation, so the host side reads DRAM, not its last-level cache), one bare
    cudaMemcpyAsync per tensor into preallocated device tensors, nothing else on the GPU; wall clock like the e2e run."""
    dst = [tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in hs) for hs in host_sets]
    nbytes = sum(t.numel() * t.element_size() for t in host_sets[0])
    for i in range(2):
        for d, h in zip(dst[i % len(dst)], host_sets[i % len(host_sets)]):
            d.copy_(h, non_blocking=True)
    barrier(world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        for d, h in zip(dst[i % len(dst)], host_sets[i % len(host_sets)]):
            d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    return nbytes * K / (time.perf_counter() - t0) / 1e9


def run_e2e(rsm, wl, dev, world, rank, K, dtype):
    """The step through the public streaming API from PINNED HOST buffers: H2D, kernels, D2H inside the region."""
    host_sets = [tuple(t.pin_memory() for t in wl.host_inputs(99 + 17 * rank + s, dtype=dtype)) for s in range(2)]
    out_dtype = dtype
    outs = [torch.empty((wl.N, wl.H, wl.W), dtype=out_dtype).pin_memory() for _ in range(3)]
    h2d = sum(t.numel() * t.element_size() for t in host_sets[0])
    d2h = outs[0].numel() * outs[0].element_size()

    # serial form: copy, compute, copy back, wait (what a caller without the pipeline gets)
    def e2e_serial(i):
        inp = tuple(t.to(dev, non_blocking=True) for t in host_sets[i % 2])
        disp = wl.step(rsm, inp)[2]
        outs[0].copy_(disp, non_blocking=True)
        torch.cuda.synchronize()

    for i in range(3):
        e2e_serial(i)
    barrier(world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        e2e_serial(i)
    torch.cuda.synchronize()
    serial_s = max_over_ranks(time.perf_counter() - t0, world, dev)

    pipe = rsm.HostPipeline(lambda l, r, c: wl.step(rsm, (l, r, c))[2], device=dev, depth=2)
    for _ in pipe.run((host_sets[i % 2] for i in range(4)), outs):
        pass
    barrier(world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_done = sum(1 for _ in pipe.run((host_sets[i % 2] for i in range(K)), outs))
    torch.cuda.synchronize()
    mine = time.perf_counter() - t0
    e2e_s = max_over_ranks(mine, world, dev)
    assert n_done == K
    return {"seconds": e2e_s, "serial_seconds": serial_s, "h2d": h2d, "d2h": d2h, "h2d_GBps_this_rank": h2d * K / mine / 1e9,
            "host_sets": host_sets}


def run_b200(args, wl):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    import realtime_stereo_matcher_b200 as rsm
    rsm.load_library()
    world, rank, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    peak, peak_src = measured_peak()

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

        # ---- a >= 2 s back-to-back run of the dominant kernel (clocks sampled across it as well)
        sustained = sustained_concat(rsm, wl, sets, peak) if not args.quick else None
        barrier(world)

        # ---- end to end through the public API with HOST buffers (pinned), copies inside the region
        e32 = run_e2e(rsm, wl, dev, world, rank, K, torch.float32)
        e16 = run_e2e(rsm, wl, dev, world, rank, K, torch.float16)
        barrier(world)
        clocks = sampler.stop() if rank == 0 else None   # sampled across the device-timed, sustained and e2e regions
        ceiling = h2d_ceiling(dev, e32["host_sets"], K, world)
        ceiling_all = max_over_ranks(-ceiling, world, dev)   # min over ranks
        extra = tensor_core_kernels(rsm, dev, peak) if (rank == 0 and not args.quick) else {}
        barrier(world)

    pairs = wl.pairs_per_step * world * K
    value = pairs / (ms_total * 1e-3)
    if rank != 0:
        return
    # ---- CPU baseline beside it (rank 0, N=1 only): the reference's CPU path on the same per-GPU batch
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, info = time_cpu_reference(wl)
        cpu = {"value": v, "unit": "pairs/s", "cores": info["cores"], "kind": info["kind"],
               "sample": f"all {wl.pairs_per_step} pairs of one step, best of {info['reps']} passes "
                         f"({info['best_s']:.2f} s each), " +
                         ("the unmodified reference from baseline/_ref" if info["kind"] == "reference"
                          else "torch CPU port of the reference op sequence")}
    ab = wl.algorithmic_bytes()
    names = list(ab)
    dom = wl.dominant
    achieved = ab[names[dom]] / (op_ms[dom] * 1e-3) / 1e9
    kernels = {n: {"ms": op_ms[j], "algorithmic_GBps": ab[n] / (op_ms[j] * 1e-3) / 1e9,
                   "frac_of_hbm_peak": ab[n] / (op_ms[j] * 1e-3) / 1e9 / peak} for j, n in enumerate(names)}
    if "v4_head" in kernels and getattr(wl, "D", 0):
        # the head is bound by the MUFU / FMA pipes, not by HBM: two ex2 per four fine disparities at
        # 16 ex2 / clk / SM is its floor on this GPU
        fine = wl.pairs_per_step * wl.D * wl.H * wl.W
        floor_ms = fine / 2 / (148 * 16 * 1.965e9) * 1e3
        kernels["v4_head"].update({"bound": "mufu/fma pipes (not hbm)", "G_fine_disparities_per_s": fine / (kernels["v4_head"]["ms"] * 1e-3) / 1e9,
                                   "mufu_floor_ms": floor_ms, "frac_of_mufu_floor": floor_ms / kernels["v4_head"]["ms"]})
    kernels.update(extra)
    line = {
        "metric": "stereo_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": K,
        "warmup": Wm, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
        "config": config_dict(wl, world),
        "roofline": {"bound": "hbm", "kernel": wl.kernels[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(wl.kernels[dom]), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ab[names[dom]], "avg_launch_ms": op_ms[dom],
                     "sustained": sustained},
        "kernels": kernels,
        "cpu_baseline": cpu,
        "e2e": {"value": pairs / e32["seconds"], "unit": "pairs/s", "h2d_bytes_per_step": e32["h2d"],
                "d2h_bytes_per_step": e32["d2h"], "serial_value": pairs / e32["serial_seconds"],
                "h2d_GBps_rank0": e32["h2d_GBps_this_rank"], "h2d_ceiling_GBps_min_rank": -ceiling_all,
                "f16_features": {"value": pairs / e16["seconds"], "h2d_bytes_per_step": e16["h2d"],
                                 "d2h_bytes_per_step": e16["d2h"], "h2d_GBps_rank0": e16["h2d_GBps_this_rank"],
                                 "note": "same step with fp16 host features / cost (what autocast evaluation feeds "
                                         "the path): half the bytes on the PCIe link, kernels run in fp16"},
                "note": "public API HostPipeline: per step pinned host features+cost -> H2D (copy-in stream) -> 3 "
                        "kernels -> D2H of the disparity map (copy-out stream), device buffers allocated once; "
                        "serial_value: no overlap; h2d_ceiling: bare cudaMemcpyAsync of the same pinned buffers, every rank at once, on "
                        "this box; the volumes stay in HBM for the aggregation network, as in the model"},
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
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS) + ["cfg5_train"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the sustained loop and the extra tensor-core kernels")
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: everything written to fd 1 from here on (NCCL's version
    # banner under NCCL_DEBUG, library chatter) is sent to stderr, the result goes to a private copy of stdout
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.workload == "cfg5_train":
        import importlib.util
        spec = importlib.util.spec_from_file_location("rsm_train_bench", os.path.join(ROOT, "tools", "train_bench.py"))
        train_bench = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(train_bench)
        train_bench.run(args, _RESULT_OUT)
    else:
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
