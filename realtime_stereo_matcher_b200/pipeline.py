"""Host-side streaming of stereo batches: pinned host buffers -> device -> hot path -> host.

The hot-path kernels are HBM-bound at ~1 ms per 8-pair batch, far faster than PCIe can feed them
(107 MB of fp32 features + cost per batch), so an inference service overlaps the host->device copy
of batch i+1 with the kernels of batch i and the device->host copy of result i-1.  ``HostPipeline``
is that ring: a copy-in stream, the caller's stream for compute, a copy-out stream, events for the
hand-offs, device input buffers allocated ONCE (``depth`` sets; no allocator traffic and no
``record_stream`` bookkeeping in the steady state), no host synchronisation except when a result is
consumed.  One instance per GPU / rank; batches are independent, so there is no cross-rank traffic
(SURVEY.md 8e).  Host tensors may be fp32 or 16-bit: the kernels take fp16 / bf16 features as they
are (what autocast evaluation produces), which halves the bytes on the PCIe link.
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch


class HostPipeline:
    def __init__(self, step: Callable[..., torch.Tensor], device=None, depth: int = 2):
        """``step(*device_tensors) -> device tensor`` is the per-batch work (e.g. volumes + v4 head)."""
        if not torch.cuda.is_available():
            raise RuntimeError("HostPipeline needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.step = step
        self.depth = max(1, depth)
        self.copy_in = torch.cuda.Stream(self.device)
        self.copy_out = torch.cuda.Stream(self.device)
        self._slots: Optional[List[Tuple[torch.Tensor, ...]]] = None
        self._slot_free: List[Optional[torch.cuda.Event]] = []
        self.h2d_bytes = 0          # bytes copied host -> device / device -> host since construction
        self.d2h_bytes = 0

    def _ensure_slots(self, batch: Sequence[torch.Tensor]) -> None:
        sig = [(tuple(t.shape), t.dtype) for t in batch]
        if self._slots is not None and [(tuple(t.shape), t.dtype) for t in self._slots[0]] == sig:
            return
        self._slots = [tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in batch)
                       for _ in range(self.depth + 1)]
        self._slot_free = [None] * (self.depth + 1)

    def run(self, host_batches: Iterable[Sequence[torch.Tensor]], host_out: Sequence[torch.Tensor]) -> Iterator[Tuple[int, torch.Tensor]]:
        """Yield ``(i, host_out[i % len(host_out)])`` once batch i's result has landed on the host.

        ``host_batches`` yields tuples of PINNED host tensors; ``host_out`` is a ring of pinned result
        buffers (depth + 1 of them for full overlap; fewer only shortens the pipeline).  Per batch the region covers its H2D copy, the kernels and the D2H copy."""
        compute = torch.cuda.current_stream(self.device)
        inflight = []   # (index, done_event, out_buffer)
        limit = max(1, min(self.depth + 1, len(host_out)))      # a result buffer is rewritten only after it was yielded
        for i, batch in enumerate(host_batches):
            while len(inflight) >= limit:
                j, ev, buf = inflight.pop(0)
                ev.synchronize()
                yield j, buf
            self._ensure_slots(batch)
            k = i % len(self._slots)
            dev_in = self._slots[k]
            with torch.cuda.stream(self.copy_in):
                if self._slot_free[k] is not None:
                    self.copy_in.wait_event(self._slot_free[k])      # the kernels that read this slot have finished
                for d, h in zip(dev_in, batch):
                    d.copy_(h, non_blocking=True)                    # one cudaMemcpyAsync per tensor
                    self.h2d_bytes += h.numel() * h.element_size()
                copied = torch.cuda.Event()
                copied.record(self.copy_in)
            compute.wait_event(copied)
            res = self.step(*dev_in)
            computed = torch.cuda.Event()
            computed.record(compute)
            self._slot_free[k] = computed
            out = host_out[i % len(host_out)]
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(computed)
                out.copy_(res, non_blocking=True)
                res.record_stream(self.copy_out)
                self.d2h_bytes += out.numel() * out.element_size()
                done = torch.cuda.Event()
                done.record(self.copy_out)
            inflight.append((i, done, out))
        for j, ev, buf in inflight:
            ev.synchronize()
            yield j, buf
