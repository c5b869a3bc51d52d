"""Multi-GPU layout of the path: stereo pairs are independent, so the batch is
sharded across ranks (one process per GPU) with NO collective on the data path.  NCCL is used
only outside it: DDP gradient all-reduce in training and the gather of scalar metrics here."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of ``n_items`` stereo pairs owned by ``rank``; the first
    ``n_items % world_size`` ranks take one extra pair, so shards differ by at most one."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def all_gather_metrics(metrics: Dict[str, float], device=None) -> Dict[str, list]:
    """Gather a dict of per-rank scalars to every rank (key order = sorted keys).  Works with the
    nccl backend (tensors on ``device``) and with gloo on CPU; without an initialised process group
    it returns single-element lists."""
    keys = sorted(metrics)
    if not (dist.is_available() and dist.is_initialized()):
        return {k: [float(metrics[k])] for k in keys}
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.tensor([float(metrics[k]) for k in keys], dtype=torch.float64, device=device)
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return {k: [float(t[i]) for t in out] for i, k in enumerate(keys)}
