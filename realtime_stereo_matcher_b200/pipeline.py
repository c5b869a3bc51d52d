"""Host-side streaming of stereo batches: pinned host buffers -> device -> hot path -> host.

The hot-path kernels are HBM-bound at ~1 ms per 8-pair batch, far faster than PCIe can feed them
(107 MB of fp32 features + cost per batch), so an inference service overlaps the host->device copy
of batch i+1 with the kernels of batch i.  ``HostPipeline`` is that double buffer: a copy stream
for H2D, the caller's stream for compute and the D2H of the (small) result, events for the
hand-offs, no host synchronisation except when a result is consumed.  One instance per GPU / rank;
batches are independent, so there is no cross-rank traffic.
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, Sequence, Tuple

import torch


class HostPipeline:
    def __init__(self, step: Callable[..., torch.Tensor], device=None, depth: int = 2):
        """``step(*device_tensors) -> device tensor`` is the per-batch work (e.g. volumes + v4 head)."""
        if not torch.cuda.is_available():
            raise RuntimeError("HostPipeline needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.step = step
        self.depth = max(1, depth)
        self.copy_stream = torch.cuda.Stream(self.device)

    def run(self, host_batches: Iterable[Sequence[torch.Tensor]], host_out: Sequence[torch.Tensor]) -> Iterator[Tuple[int, torch.Tensor]]:
        """Yield ``(i, host_out[i % len(host_out)])`` once batch i's result has landed on the host.

        ``host_batches`` yields tuples of PINNED host tensors; ``host_out`` is a ring of pinned result
        buffers (len >= depth).  Per batch the region covers its H2D copy, the kernels and the D2H copy."""
        compute = torch.cuda.current_stream(self.device)
        inflight = []   # (index, done_event, out_buffer)
        for i, batch in enumerate(host_batches):
            with torch.cuda.stream(self.copy_stream):
                dev_in = tuple(t.to(self.device, non_blocking=True) for t in batch)
                copied = torch.cuda.Event()
                copied.record(self.copy_stream)
            compute.wait_event(copied)
            res = self.step(*dev_in)
            for t in dev_in:                     # the copy stream allocated them; compute uses them
                t.record_stream(compute)
            out = host_out[i % len(host_out)]
            out.copy_(res, non_blocking=True)
            done = torch.cuda.Event()
            done.record(compute)
            inflight.append((i, done, out))
            if len(inflight) >= self.depth:
                j, ev, buf = inflight.pop(0)
                ev.synchronize()
                yield j, buf
        for j, ev, buf in inflight:
            ev.synchronize()
            yield j, buf
