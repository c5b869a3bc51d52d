"""bench.py --workload cfg5_train: the training half of the multi-GPU story (BASELINE config 5: "full-res 1080p
soft-argmin + argmin regression at D=192 streaming batch, forward+backward for train_stereo at 2/4/8 GPUs";
reference loop train_stereo.py:138-212).

One step per rank (B pairs of 1080x1920):
  (a) the regression kernels alone on a resident (B,192,1080,1920) cost: soft-argmax + hard argmin in one pass
      (rsm_regress_fwd), smooth-L1 against a target, backward through rsm_regress_bwd;
  (b) a small DispNetC/v4-style network wrapped in DistributedDataParallel whose hot path runs on the kernels:
      conv features at 1/4 res -> mean-correlation volume D=48 (rsm_inner_fwd/bwd) -> conv aggregation -> the v4 head
      to full resolution, D=192 (rsm_upsample_regress_fwd/bwd) -> SequenceLoss (rsm_seqloss_fwd/bwd) -> backward with
      the NCCL gradient all-reduce inside the timed region -> AdamW step.
Pairs shard across ranks (weak scaling); the only collective is DDP's all-reduce.  Reported: pairs/s, ms/step, the
exposed all-reduce share (same steps under no_sync()), a bare all-reduce of the gradient bucket, e2e from pinned
host images.  `--impl reference` runs the same step with the reference's own ops on the CPU (1 pair per step).
"""
from __future__ import annotations

import json
import os
import statistics
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NAME = ("cfg5_train: 2 pairs/GPU @1080x1920 -- regression fwd+bwd on (2,192,1080,1920) + DDP step of a small stereo net "
        "(corr D=48 at 1/4 res -> v4 head D=192 -> SequenceLoss -> backward + all-reduce -> AdamW)")
B, D, H, W, D4 = 2, 192, 1080, 1920, 48


class TrainStereo(nn.Module):
    """conv features (1/4 res) -> mean-correlation volume -> conv aggregation -> v4 head at full resolution.
    `ops` supplies make_correlation_volume(l, r, d) and v4_head(cost, D, H, W)."""

    def __init__(self, ops, c=16):
        super().__init__()
        self.ops = ops
        self.feat = nn.Sequential(nn.Conv2d(3, c, 5, 2, 2), nn.ReLU(), nn.Conv2d(c, c, 3, 2, 1))
        self.agg = nn.Sequential(nn.Conv2d(D4, D4, 3, 1, 1), nn.ReLU(), nn.Conv2d(D4, D4, 3, 1, 1))

    def forward(self, left, right):
        lf, rf = self.feat(left / 255.0), self.feat(right / 255.0)
        vol = self.ops.make_correlation_volume(lf, rf, D4)
        pred = self.ops.v4_head(self.agg(vol), D, left.shape[2], left.shape[3])
        return [-1.0 * pred.unsqueeze(1)]


class RsmOps:
    def __init__(self):
        import realtime_stereo_matcher_b200 as rsm
        self.rsm = rsm
        self.make_correlation_volume = rsm.make_correlation_volume
        self.v4_head = rsm.v4_head
        self.loss = rsm.SequenceLoss(0.9, 700.0)

    def regress(self, cost):
        soft, amin, _ = self.rsm.regress(cost, argmin=True, argmax=False)
        return soft, amin


class RefOps:
    """The reference's own functions (baseline/_ref) -- CPU arm."""

    def __init__(self):
        from oracle import ref_loader
        ref = ref_loader.load()
        self.ref = ref
        import importlib
        self.make_correlation_volume = importlib.import_module("model.mobile_disp_net_c").make_correlation_volume
        self.loss = ref.loss_loss.SequenceLoss(0.9, 700.0)

    def v4_head(self, cost, d, h, w):          # model/mobile_stereo_net_v4.py:511-518
        c = F.interpolate(cost.unsqueeze(1), [d, h, w], mode="trilinear").squeeze(1)
        return self.ref.v4.disparity_regression(F.softmax(c, dim=1), d)

    def regress(self, cost):                   # model/mobile_stereo_net.py:144-147 + torch.argmin
        p = F.softmax(cost, dim=1)
        dv = torch.arange(0, cost.shape[1], dtype=cost.dtype, device=cost.device).view(1, -1, 1, 1)
        return torch.sum(p * dv, 1), torch.argmin(cost.detach(), dim=1)


def host_batch(seed, b):
    g = torch.Generator().manual_seed(seed)
    left = torch.rand((b, 3, H, W), generator=g) * 255.0
    right = torch.roll(left, -7, 3)
    gt = -(7.0 + torch.rand((b, 1, H, W), generator=g))
    valid = (torch.rand((b, H, W), generator=g) > 0.1).float()
    return left, right, gt, valid


def train_step(ops, net, opt, cost, target, batch):
    left, right, gt, valid = batch
    opt.zero_grad(set_to_none=True)
    cost.grad = None
    soft, amin = ops.regress(cost)
    l_reg = F.smooth_l1_loss(soft, target)
    l_reg.backward()
    preds = net(left, right)
    loss = ops.loss(preds, gt, valid)
    loss.backward()                      # DDP: gradient all-reduce overlaps the tail of this
    opt.step()
    return loss.detach(), amin


def config_dict(world):
    return {"workload": NAME, "pairs_per_gpu_per_step": B, "l2": "each step streams 3.2 GB of cost + 3.2 GB of gradient (>> L2)",
            "parallelism": f"dp{world} (batch sharded; DDP NCCL all-reduce of the gradients)"}


def run_reference(args, out):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(1234)
    ops = RefOps()
    net = TrainStereo(ops)
    opt = torch.optim.AdamW(net.parameters(), lr=2e-4)
    b = 1
    cost = (torch.randn((b, D, H, W)) * 3).requires_grad_(True)
    target = torch.rand((b, H, W)) * 100
    batch = host_batch(7, b)
    for _ in range(args.warmup):
        train_step(ops, net, opt, cost, target, batch)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        train_step(ops, net, opt, cost, target, batch)
    dt = time.perf_counter() - t0
    value = b * args.steps / dt
    desc = (f"1 of {B} pairs per step (same shapes), the reference's own ops from baseline/_ref (make_correlation_volume, "
            f"F.interpolate->softmax->disparity_regression, SequenceLoss) + torch autograd on the CPU, {threads} threads")
    print(json.dumps({
        "impl": "reference", "metric": "stereo_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(max(args.gpus, int(os.environ.get("WORLD_SIZE", "1")))),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "reference", "sample": desc},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
    }), file=out, flush=True)


def run(args, out):
    if args.impl == "reference":
        return run_reference(args, out)
    import bench
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    world, rank, local = bench.dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    ops = RsmOps()
    ops.rsm.load_library()
    peak, peak_src = bench.measured_peak()

    torch.manual_seed(1234)                       # same weights on every rank
    net = TrainStereo(ops).to(dev)
    model = DDP(net, device_ids=[local]) if world > 1 else net
    opt = torch.optim.AdamW(net.parameters(), lr=2e-4, weight_decay=1e-5)
    nparam = sum(p.numel() for p in net.parameters())
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    costs = [(torch.randn((B, D, H, W), device=dev, generator=g) * 3).requires_grad_(True) for _ in range(2)]
    target = torch.rand((B, H, W), device=dev, generator=g) * 100
    batches = [tuple(t.to(dev) for t in host_batch(7 + 13 * rank + s, B)) for s in range(2)]
    K, Wm = args.steps, max(args.warmup, 3)

    def steps(n, sync=True):
        import contextlib
        ctx = contextlib.nullcontext() if (sync or world == 1) else model.no_sync()
        with ctx:
            for i in range(n):
                loss, amin = train_step(ops, model, opt, costs[i % 2], target, batches[i % 2])
        return loss

    steps(Wm)
    torch.cuda.synchronize()
    sampler = bench.ClockSampler(local)
    if rank == 0:
        sampler.start()

    def timed(sync):
        bench.barrier(world)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = steps(K, sync)
        b.record()
        torch.cuda.synchronize()
        bench.barrier(world)
        return bench.max_over_ranks(a.elapsed_time(b), world, dev), float(loss)

    ms_total, last_loss = timed(True)
    ms_nosync, _ = timed(False) if world > 1 else (ms_total, None)

    # the regression kernels alone (roofline of the dominant, HBM-bound pair)
    def reg_times():
        f, bw = [], []
        for i in range(6):
            c = costs[i % 2]
            c.grad = None
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            soft, amin = ops.regress(c)
            e[1].record()
            soft.backward(torch.ones_like(soft))
            e[2].record()
            torch.cuda.synchronize()
            f.append(e[0].elapsed_time(e[1])); bw.append(e[1].elapsed_time(e[2]))
        return statistics.mean(f[1:]), statistics.mean(bw[1:])

    ms_f, ms_b = reg_times()
    bytes_f = B * D * H * W * 4 + B * H * W * (4 + 8 + 4 + 4)
    bytes_b = 2 * B * D * H * W * 4 + 3 * B * H * W * 4

    # bare all-reduce of one gradient-sized bucket
    ar_ms = None
    if world > 1:
        bucket = torch.zeros(nparam, device=dev)
        for _ in range(3):
            dist.all_reduce(bucket)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            dist.all_reduce(bucket)
        b.record()
        torch.cuda.synchronize()
        ar_ms = bench.max_over_ranks(a.elapsed_time(b) / 20, world, dev)

    # end to end: pinned host images + ground truth -> H2D -> step -> loss scalar back on the host
    hb = [tuple(t.pin_memory() for t in host_batch(50 + rank + s, B)) for s in range(2)]
    h2d = sum(t.numel() * t.element_size() for t in hb[0])
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()

    def e2e(n):
        for i in range(n):
            batch = tuple(t.to(dev, non_blocking=True) for t in hb[i % 2])
            loss, _ = train_step(ops, model, opt, costs[i % 2], target, batch)
            host_loss.copy_(loss, non_blocking=True)
            torch.cuda.synchronize()

    e2e(2)
    bench.barrier(world)
    t0 = time.perf_counter()
    e2e(K)
    e2e_s = bench.max_over_ranks(time.perf_counter() - t0, world, dev)
    clocks = sampler.stop() if rank == 0 else None
    losses = ops.rsm.all_gather_metrics({"loss": last_loss}, device=dev)

    if rank != 0:
        return
    pairs = B * world * K
    cpu = None
    line = {
        "metric": "stereo_pairs_per_sec", "value": pairs / (ms_total * 1e-3), "unit": "pairs/s", "n_gpus": world, "steps": K,
        "warmup": Wm, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_dict(world),
        "roofline": {"bound": "hbm", "kernel": "regress_bwd_kernel", "achieved": bytes_b / (ms_b * 1e-3) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": bytes_b / (ms_b * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bytes_b, "avg_launch_ms": ms_b},
        "kernels": {"regress_fwd[soft+argmin]": {"ms": ms_f, "algorithmic_GBps": bytes_f / (ms_f * 1e-3) / 1e9,
                                                  "frac_of_hbm_peak": bytes_f / (ms_f * 1e-3) / 1e9 / peak},
                    "regress_bwd(+ones)": {"ms": ms_b, "algorithmic_GBps": bytes_b / (ms_b * 1e-3) / 1e9,
                                           "frac_of_hbm_peak": bytes_b / (ms_b * 1e-3) / 1e9 / peak}},
        "ddp": {"params": nparam, "ms_per_step_no_sync": ms_nosync / K,
                "exposed_allreduce_ms_per_step": (ms_total - ms_nosync) / K,
                "exposed_allreduce_share": (ms_total - ms_nosync) / ms_total,
                "bare_allreduce_ms": ar_ms, "per_rank_loss": losses["loss"]},
        "cpu_baseline": cpu,
        "e2e": {"value": pairs / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "note": "pinned host images + ground truth + mask -> H2D -> the whole training step -> loss scalar D2H"},
        "gpu_launches": K * 12,
        "clocks": clocks,
    }
    print(json.dumps(line), file=out, flush=True)
