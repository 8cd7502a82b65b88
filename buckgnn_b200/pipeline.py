"""Double-buffered host->device staging for inference loops.

The reference scripts do `batch = batch.to(device)` inside the loop
(`INFERENCE.py:135`, `INFERENCE_TIMER.py:231`), i.e. the copy of batch i+1 waits for the
forward of batch i.  `DevicePrefetcher` keeps two device-side staging batches and issues the
H2D copies of the next pinned batch on a separate CUDA stream while the current forward
runs, so a B200 forward (about 12 ms for 256 plates) hides the ~5 ms PCIe transfer.
"""
from __future__ import annotations

from collections import deque
from typing import Callable, Iterable, Iterator, Optional

import torch

from dataclasses import dataclass

from . import capi
from .synth import PlateBatch


@dataclass
class WireBatch:
    """Host-side wire format of an inference batch for models that never read `edge_attr` (every GraphSAGE variant):
    what has to cross PCIe, and nothing else.

    * `x` [N, F] f32 -- as in the PyG batch;
    * `edges` [2, E_wire] int32 -- the explicit edges only, global node ids (a batch has < 2^31 nodes).  The super
      node's hub pairs (s, i), (i, s) -- a third of all directed edges -- are implicit: `create_super_node` appends
      them after a graph's mesh edges in a fixed order (reference Dataset_Preparation/VirtualEdgeCreate.py:106-111), so
      the device rebuilds them from the node offsets;
    * `node_ptr`, `wire_ptr`, `full_ptr` [G+1] int64 -- node offsets, offsets into `edges`, offsets into the full
      PyG `edge_index`; the `batch` vector is rebuilt from `node_ptr`.
    `edge_attr`, `y` and `ptr` stay on the host.  `bg_expand_wire` turns this into the exact tensors
    `batch.to(device)` would have produced (tests/test_gpu_kernels.py), 98 MB instead of 294 MB for 256 plates."""
    x: torch.Tensor
    edges: torch.Tensor
    node_ptr: torch.Tensor
    wire_ptr: torch.Tensor
    full_ptr: torch.Tensor
    num_graphs: int
    num_nodes: int
    num_edges: int            # of the full PyG edge_index
    edge_features: int

    FIELDS = ("x", "edges", "node_ptr", "wire_ptr", "full_ptr")

    @staticmethod
    def from_batch(b: PlateBatch) -> "WireBatch":
        """Host-side (loader) conversion.  A graph's trailing edges are dropped from the wire only if they are EXACTLY
        the hub pairs of its last node in the reference's order; any other graph ships all its edges."""
        ei = b.edge_index
        g = b.num_graphs
        node_ptr = b.ptr.to(torch.int64)
        E = ei.shape[1]
        # edges are grouped by graph (PyG collate): graph of an edge = graph of its source node
        eg = b.batch[ei[0]] if E > 0 else torch.zeros(0, dtype=torch.int64)
        counts = torch.bincount(eg, minlength=g)
        full_ptr = torch.zeros(g + 1, dtype=torch.int64)
        full_ptr[1:] = torch.cumsum(counts, 0)
        keep = torch.ones(E, dtype=torch.bool)
        # the per-graph edge offsets above mean something only if the edge list IS grouped by graph; a batch built any
        # other way ships every edge (wire_ptr == full_ptr: bg_expand_wire is then a plain widening copy)
        grouped = E == 0 or bool((eg[1:] >= eg[:-1]).all())
        for k in range(g if grouped else 0):
            n0, n1 = int(node_ptr[k]), int(node_ptr[k + 1])
            f0, f1 = int(full_ptr[k]), int(full_ptr[k + 1])
            m = n1 - n0 - 1                                    # hub pairs if the last node is a super node
            if m <= 0 or f1 - f0 < 2 * m:
                continue
            tail = ei[:, f1 - 2 * m:f1]
            s = n1 - 1
            others = torch.arange(n0, s)
            if (torch.equal(tail[0, 0::2], torch.full((m,), s)) and torch.equal(tail[1, 0::2], others) and
                    torch.equal(tail[0, 1::2], others) and torch.equal(tail[1, 1::2], torch.full((m,), s))):
                keep[f1 - 2 * m:f1] = False
        edges = ei[:, keep].to(torch.int32).contiguous()
        wire_counts = torch.bincount(eg[keep], minlength=g)
        wire_ptr = torch.zeros(g + 1, dtype=torch.int64)
        wire_ptr[1:] = torch.cumsum(wire_counts, 0)
        return WireBatch(b.x.to(torch.float32).contiguous(), edges, node_ptr.contiguous(), wire_ptr, full_ptr, g,
                         b.num_nodes, E, b.edge_attr.shape[1])

    def pin_memory(self) -> "WireBatch":
        return WireBatch(*[getattr(self, f).pin_memory() for f in self.FIELDS], self.num_graphs, self.num_nodes,
                         self.num_edges, self.edge_features)

    def nbytes(self) -> int:
        return sum(getattr(self, f).numel() * getattr(self, f).element_size() for f in self.FIELDS)

    def expand(self, device, out: Optional[PlateBatch] = None, stream=None) -> PlateBatch:
        """Device-resident wire tensors (self must live on `device`) -> PyG-layout batch (edge_attr: a 0-row
        placeholder of the right width, `y` / `ptr` as given by node_ptr)."""
        dev = torch.device(device)
        if out is None or out.edge_index.shape[1] != self.num_edges or out.batch.shape[0] != self.num_nodes:
            out = PlateBatch(self.x, torch.empty((2, self.num_edges), dtype=torch.int64, device=dev),
                             torch.empty((0, self.edge_features), dtype=torch.float32, device=dev),
                             torch.empty(self.num_nodes, dtype=torch.int64, device=dev),
                             torch.empty(0, dtype=torch.float32, device=dev), self.node_ptr, self.num_graphs)
        out.x, out.ptr, out.num_graphs = self.x, self.node_ptr, self.num_graphs
        s = (stream or torch.cuda.current_stream(dev)).cuda_stream
        capi.expand_wire(self.edges.data_ptr(), self.edges.shape[1], self.node_ptr.data_ptr(), self.wire_ptr.data_ptr(),
                         self.full_ptr.data_ptr(), self.num_graphs, self.num_edges, out.edge_index.data_ptr(),
                         out.batch.data_ptr(), s)
        return out


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
        # x / edge_index / batch go first and get their own event: a model that never reads edge_attr (the GraphSAGE
        # variants) can start on a batch while its edge_attr / y / ptr are still crossing PCIe
        self.core_copied = [torch.cuda.Event(), torch.cuda.Event()]
        self.wait_for_all_fields = True
        self.released = [torch.cuda.Event(), torch.cuda.Event()]
        self.copy_events = []          # (start, stop) per staged batch when `time_copies` is set
        self.time_copies = False
        # bg_expand_wire on the CONSUMER's stream, right before its forward (default), instead of on the copy stream:
        # a kernel that lands beside the forward's persistent kernels (one CTA per SM, rows split statically) holds
        # up the CTAs of the SMs it occupies and with them the whole kernel -- measured 0.37 ms per forward against
        # the ~0.05 ms the expansion costs in line
        self.expand_on_compute = True

    def _stage_wire(self, host: "WireBatch", k: int) -> PlateBatch:
        """WireBatch: copy its five tensors, then rebuild edge_index / batch on the copy stream (bg_expand_wire)."""
        slot = self.slots[k]
        fits = (isinstance(slot, tuple) and all(getattr(slot[0], f).shape == getattr(host, f).shape for f in WireBatch.FIELDS)
                and slot[1].edge_index.shape[1] == host.num_edges)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.released[k])
            if not fits:
                wire = WireBatch(*[torch.empty_like(getattr(host, f), device=self.device) for f in WireBatch.FIELDS],
                                 host.num_graphs, host.num_nodes, host.num_edges, host.edge_features)
                slot = (wire, None)
            wire, full = slot
            wire.num_graphs, wire.num_nodes, wire.num_edges = host.num_graphs, host.num_nodes, host.num_edges
            if self.time_copies:
                t0 = torch.cuda.Event(enable_timing=True)
                t0.record(self.copy_stream)
            for f in WireBatch.FIELDS:
                getattr(wire, f).copy_(getattr(host, f), non_blocking=True)
            if self.time_copies:
                t1 = torch.cuda.Event(enable_timing=True)
                t1.record(self.copy_stream)
                self.copy_events.append((t0, t1))
            if not self.expand_on_compute:
                full = wire.expand(self.device, full, stream=self.copy_stream)
            self.slots[k] = (wire, full)
            self.core_copied[k].record(self.copy_stream)
            self.copied[k].record(self.copy_stream)
        return full if not self.expand_on_compute else wire

    def _stage(self, host, k: int) -> PlateBatch:
        if isinstance(host, WireBatch):
            return self._stage_wire(host, k)
        slot = self.slots[k]
        if isinstance(slot, tuple):
            slot = None
        fields = ("edge_index", "batch", "x", "edge_attr", "y", "ptr")
        fits = slot is not None and all(getattr(slot, f).shape == getattr(host, f).shape for f in fields)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.released[k])       # the forward that used this slot is done
            if not fits:
                slot = PlateBatch(*[torch.empty_like(getattr(host, f), device=self.device)
                                    for f in ("x", "edge_index", "edge_attr", "batch", "y", "ptr")], host.num_graphs)
                self.slots[k] = slot
            slot.num_graphs = host.num_graphs
            if self.time_copies:
                t0 = torch.cuda.Event(enable_timing=True)
                t0.record(self.copy_stream)
            for f in fields:
                getattr(slot, f).copy_(getattr(host, f), non_blocking=True)
                if f == "x":
                    self.core_copied[k].record(self.copy_stream)
            if self.time_copies:
                t1 = torch.cuda.Event(enable_timing=True)
                t1.record(self.copy_stream)
                self.copy_events.append((t0, t1))
            self.copied[k].record(self.copy_stream)
        return slot

    def __iter__(self) -> Iterator[PlateBatch]:
        return self.iterate(self.batches)

    def iterate(self, batches: Iterable[PlateBatch]) -> Iterator[PlateBatch]:
        """One pass over `batches`; the two device-side staging batches persist between passes (a long-lived
        loader allocates them once)."""
        it = iter(batches)
        k = 0
        try:
            nxt = self._stage(next(it), k)
        except StopIteration:
            return
        while nxt is not None:
            cur, cur_k = nxt, k
            torch.cuda.current_stream(self.device).wait_event(
                self.copied[cur_k] if self.wait_for_all_fields else self.core_copied[cur_k])
            if isinstance(cur, WireBatch):                      # expand in line, on the consumer's stream
                wire, full = self.slots[cur_k]
                cur = wire.expand(self.device, full)
                self.slots[cur_k] = (wire, cur)
            k ^= 1
            try:
                nxt = self._stage(next(it), k)                  # overlaps with the caller's forward on `cur`
            except StopIteration:
                nxt = None
            yield cur
            self.released[cur_k].record(torch.cuda.current_stream(self.device))


class PipelinedInference:
    """The inference loop of `INFERENCE.py:133-150` / `INFERENCE_TIMER.py:229-237` as a three-stage pipeline:
    while batch i computes, batch i+1 is copied host -> device on a second stream (`DevicePrefetcher`) and the
    eigenvalues of batch i-1 are already on their way to pinned host memory.  The GPU never waits for the host to
    read a result before it gets the next batch's kernels, so a step costs max(device time, host enqueue time)
    instead of their sum.

        for step, pred_host in PipelinedInference(model, pinned_batches, "cuda:0"):
            ...                                   # pred_host: [G] f32 CPU tensor of batch number `step`

    Results come back `depth` steps late, in order, every batch exactly once (the device-side staging buffers of a
    batch are recycled two steps later, so only the prediction is handed out).  The read-back goes through
    `bg_publish_words` (SM stores to pinned memory) rather than a D2H memcpy, which would queue on a copy engine
    behind the next batch's H2D transfer."""

    def __init__(self, model, batches: Iterable[PlateBatch], device, depth: int = 1,
                 on_launch: Optional[Callable] = None):
        self.model, self.batches, self.device, self.depth = model, batches, torch.device(device), max(0, int(depth))
        self.prefetcher = DevicePrefetcher(batches, self.device)
        # the GraphSAGE variants never read edge_attr (nor y / ptr): their forward may start once x, edge_index and
        # batch have arrived; the EA-GNN variants wait for the whole batch
        self.prefetcher.wait_for_all_fields = getattr(model, "model_name", "") in ("EA_GNN", "EA_GNN_Shared", "EAGNN_SAG")
        self.on_launch = on_launch            # called as on_launch(step, before: bool) around each forward (timing hooks)
        self._pool = {}                       # pinned result buffers by size

    def _read_back(self, pred: torch.Tensor, host: torch.Tensor) -> torch.cuda.Event:
        flat = pred.detach().reshape(-1)
        if flat.dtype != torch.float32:
            flat = flat.float()
        s = torch.cuda.current_stream(self.device).cuda_stream
        n = flat.numel()
        for off in range(0, n, 256):          # 4-byte words, at most 256 per call
            capi.publish_words(flat.data_ptr() + 4 * off, host.data_ptr() + 4 * off, min(256, n - off), s)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))     # the stream the publish kernels were queued on
        self._keep = flat                     # alive until the next call's kernels are queued behind it
        return ev

    def __iter__(self) -> Iterator:
        return self.run(self.batches)

    def run(self, batches: Iterable[PlateBatch]) -> Iterator:
        """One pass over `batches` (any iterable of pinned host batches).  Staging buffers and the pinned result
        buffers are kept on the object, so later passes allocate nothing."""
        inflight = deque()
        pool = self._pool
        step = 0
        with torch.no_grad():
            for b in self.prefetcher.iterate(batches):
                if self.on_launch:
                    self.on_launch(step, True)
                pred, _ = self.model(b.x, b.edge_index, b.edge_attr, b.batch)
                if self.on_launch:
                    self.on_launch(step, False)
                n = max(pred.numel(), 1)
                bufs = pool.setdefault(n, [])
                host = bufs.pop() if bufs else torch.empty(n, dtype=torch.float32).pin_memory()
                ev = self._read_back(pred, host)
                inflight.append((host, ev, tuple(pred.shape), step, n))
                step += 1
                while len(inflight) > self.depth:
                    yield self._finish(inflight.popleft(), pool)
            while inflight:
                yield self._finish(inflight.popleft(), pool)

    @staticmethod
    def _finish(item, pool):
        host, ev, shape, step, n = item
        ev.synchronize()
        count = 1
        for d in shape:
            count *= d
        result = host[:count].clone().view(shape)     # the pinned buffer goes back to the pool
        pool[n].append(host)
        return step, result
