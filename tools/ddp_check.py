#!/usr/bin/env python
"""Training-path check on 2+ GPUs (run under torchrun, NCCL): a small DispNetC-style model whose
correlation volume + soft-argmax run on the CUDA kernels, wrapped in DistributedDataParallel.
Each rank gets its shard of the batch; after backward the all-reduced gradients must equal the
gradients of a single-process run over the WHOLE batch.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_check.py
"""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import realtime_stereo_matcher_b200 as rsm  # noqa: E402
from realtime_stereo_matcher_b200.sharding import all_gather_metrics, shard_range  # noqa: E402


class TinyStereo(nn.Module):
    """conv features -> mean-correlation volume (rsm) -> conv aggregation -> soft-argmax (rsm)."""

    def __init__(self, c=16, d=24):
        super().__init__()
        self.d = d
        self.feat = nn.Sequential(nn.Conv2d(3, c, 3, 2, 1), nn.ReLU(), nn.Conv2d(c, c, 3, 1, 1))
        self.agg = nn.Conv2d(d, d, 3, 1, 1)

    def forward(self, left, right):
        lf, rf = self.feat(left), self.feat(right)
        vol = rsm.make_correlation_volume(lf, rf, self.d)
        return rsm.softmax_regression(self.agg(vol), keepdim=True)


def main():
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    torch.manual_seed(1234)
    net = TinyStereo().to(dev)
    g = torch.Generator().manual_seed(7)
    n = 4 * world
    left = torch.rand((n, 3, 64, 128), generator=g).to(dev)
    right = torch.roll(left, -3, 3)
    target = torch.rand((n, 1, 32, 64), generator=g).to(dev) * 8

    # single-process reference over the whole batch (same weights on every rank: same seed)
    ref = TinyStereo().to(dev)
    ref.load_state_dict(net.state_dict())
    (ref(left, right) - target).abs().mean().backward()

    ddp = nn.parallel.DistributedDataParallel(net, device_ids=[dev.index])
    b, e = shard_range(n, rank, world)
    loss = (ddp(left[b:e], right[b:e]) - target[b:e]).abs().mean()
    loss.backward()
    worst = 0.0
    for (name, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        worst = max(worst, float((p.grad - q.grad).abs().max() / (q.grad.abs().max() + 1e-12)))
    m = all_gather_metrics({"loss": float(loss), "grad_rel_err": worst}, device=dev)
    if rank == 0:
        print({"world": world, "per_rank_loss": m["loss"], "max_grad_rel_err": max(m["grad_rel_err"])})
        assert max(m["grad_rel_err"]) < 1e-4, m
        print("ddp_check ok: all-reduced gradients match the single-process gradients")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
