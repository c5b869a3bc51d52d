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


def bind_host_to_device(device_index: int) -> list:
    """Pin the calling process to the CPU cores NVML reports as local to GPU ``device_index`` (its NUMA node), so
    that the pinned host buffers a rank allocates afterwards are first-touched next to the GPU's PCIe root.  With one
    process per GPU and ~100 MB of host input per step and rank this keeps the ranks off each other's memory
    controllers.  Returns the CPU list applied ([] when NVML or the affinity call is unavailable: nothing changes)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return []
