"""Double-buffered host->device staging for inference loops.

The reference scripts do `batch = batch.to(device)` inside the loop
(`INFERENCE.py:135`, `INFERENCE_TIMER.py:231`), i.e. the copy of batch i+1 waits for the
forward of batch i.  `DevicePrefetcher` keeps two device-side staging batches and issues the
H2D copies of the next pinned batch on a separate CUDA stream while the current forward
runs, so a B200 forward (about 12 ms for 256 plates) hides the ~5 ms PCIe transfer.
"""
from __future__ import annotations

from typing import Iterable, Iterator

import torch

from .synth import PlateBatch


class DevicePrefetcher:
    """Iterates device-resident copies of pinned host batches, one copy ahead.

    Batches must be `PlateBatch`-like (`.x .edge_index .edge_attr .batch .y .ptr .num_graphs`)
    and pinned; shapes may differ between batches (staging buffers are re-allocated when a
    batch does not fit)."""

    def __init__(self, batches: Iterable[PlateBatch], device):
        self.batches = batches
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [None, None]
        self.copied = [torch.cuda.Event(), torch.cuda.Event()]
        self.released = [torch.cuda.Event(), torch.cuda.Event()]
        self.copy_events = []          # (start, stop) per staged batch when `time_copies` is set
        self.time_copies = False

    def _stage(self, host: PlateBatch, k: int) -> PlateBatch:
        slot = self.slots[k]
        fields = ("x", "edge_index", "edge_attr", "batch", "y", "ptr")
        fits = slot is not None and all(getattr(slot, f).shape == getattr(host, f).shape for f in fields)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.released[k])       # the forward that used this slot is done
            if not fits:
                slot = PlateBatch(*[torch.empty_like(getattr(host, f), device=self.device) for f in fields],
                                  host.num_graphs)
                self.slots[k] = slot
            slot.num_graphs = host.num_graphs
            if self.time_copies:
                t0 = torch.cuda.Event(enable_timing=True)
                t0.record(self.copy_stream)
            for f in fields:
                getattr(slot, f).copy_(getattr(host, f), non_blocking=True)
            if self.time_copies:
                t1 = torch.cuda.Event(enable_timing=True)
                t1.record(self.copy_stream)
                self.copy_events.append((t0, t1))
            self.copied[k].record(self.copy_stream)
        return slot

    def __iter__(self) -> Iterator[PlateBatch]:
        it = iter(self.batches)
        k = 0
        try:
            nxt = self._stage(next(it), k)
        except StopIteration:
            return
        while nxt is not None:
            cur, cur_k = nxt, k
            torch.cuda.current_stream(self.device).wait_event(self.copied[cur_k])
            k ^= 1
            try:
                nxt = self._stage(next(it), k)                  # overlaps with the caller's forward on `cur`
            except StopIteration:
                nxt = None
            yield cur
            self.released[cur_k].record(torch.cuda.current_stream(self.device))
