"""Graph-sharded multi-GPU inference.

Graphs of a PyG batch share no edges, eval-mode BatchNorm uses running statistics and
pooling is per graph, so the forward of a shard is exact: one process per GPU
(`torchrun`), graphs partitioned across ranks, NO collective on the data path; the only
communication is one small all_gather of the per-graph predictions.
"""
from __future__ import annotations

import os
from typing import Callable, List, Sequence

import torch
import torch.distributed as dist

from .synth import PlateGraph, collate


def graph_cost(num_nodes: int, num_edges: int) -> int:
    """Work estimate of one graph: the aggregation moves one row per edge, the update GEMM
    and the epilogue two rows' worth per node."""
    return int(num_edges) + 2 * int(num_nodes)


def partition_graphs(costs: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment of graph indices to ranks.  Deterministic
    (ties: lower index first, lower rank first); every rank's list is sorted ascending so a
    shard keeps the batch's graph order."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world_size
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        parts[r].append(i)
        load[r] += costs[i]
    return [sorted(p) for p in parts]


def batches_by_node_budget(indices: Sequence[int], num_nodes: Sequence[int], node_budget: int) -> List[List[int]]:
    """Cut a rank's shard (graph indices, in order) into batches of at most `node_budget` nodes -- the inference
    loop of configs[4], whose meshes differ 8x in size, batches by memory footprint where `INFERENCE.py:94`
    batches by graph count.  A graph larger than the budget forms a batch of its own; order is kept; every
    index appears exactly once."""
    out: List[List[int]] = []
    cur: List[int] = []
    load = 0
    for i in indices:
        n = int(num_nodes[i])
        if cur and load + n > node_budget:
            out.append(cur)
            cur, load = [], 0
        cur.append(i)
        load += n
    if cur:
        out.append(cur)
    return out


def sharded_predict(graphs: Sequence[PlateGraph], forward: Callable, device, group=None) -> torch.Tensor:
    """Run `forward(batch) -> pred [G_local]` on this rank's shard of `graphs` and return the
    predictions of ALL graphs, in the original order, on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    parts = partition_graphs([graph_cost(g.num_nodes, g.num_edges) for g in graphs], world)
    mine = parts[rank]
    if mine:
        local = forward(collate([graphs[i] for i in mine]).to(device)).reshape(-1).float()
    else:
        local = torch.empty(0, dtype=torch.float32, device=device)
    out = torch.empty(len(graphs), dtype=torch.float32, device=device)
    if world == 1:
        out[torch.tensor(mine, dtype=torch.long, device=device)] = local
        return out
    width = max(len(p) for p in parts)
    padded = torch.zeros(width, dtype=torch.float32, device=device)
    padded[:len(mine)] = local
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    for r, p in enumerate(parts):
        if p:
            out[torch.tensor(p, dtype=torch.long, device=device)] = gathered[r][:len(p)]
    return out


def allreduce_gradients(params, group=None, average: bool = True) -> int:
    """One flat all-reduce of the gradients a training step produced (config 4).

    Parameters whose `.grad` is None are skipped on every rank -- `BuckGNN` registers modules the
    selected `model_name` never uses (`sage_mlps`, `edge_encoder`, `pooling_mpl`, `batch_norm`;
    reference Models/BuckGNN.py:164,184-187), which is why DDP would need
    `find_unused_parameters=True`.  Ranks must agree on which parameters have gradients (same model,
    same branch).  Gradients are flattened into one fp32 bucket (fp64 if every gradient is fp64) (about 13 MB for GraphSage_meanAggr
    6x512), reduced with a single collective (NCCL over NVLink: latency-bound at this size) and
    scattered back.  BatchNorm batch statistics stay rank-local, as a DDP port of the reference
    would have them.  Returns the number of elements reduced."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    bucket_dtype = torch.float64 if all(g.dtype == torch.float64 for g in grads) else torch.float32
    flat = torch.cat([g.reshape(-1).to(bucket_dtype) for g in grads])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g).to(g.dtype))
        off += n
    return off


class GradSync:
    """Bucketed gradient all-reduce that can overlap the backward pass (BASELINE.json configs[3]).

    All gradients of the parameters a model trains live in ONE pre-allocated fp32 buffer, laid out in
    `train.trainable_parameters(model)` order; the backward kernels write straight into its slices (no `torch.cat`,
    no copy back).  As soon as the backward has finished the gradients of a group of parameters -- the head, then
    layer L-1 ... layer 0, then the encoder -- `launch()` records an event on the compute stream and starts the
    all-reduce (mean) of that group's slice on a communication stream, so it runs under the backward kernels of the
    earlier layers; `finish()` makes the compute stream wait for the collectives at the end of the backward (only the
    last, smallest bucket is exposed).  Per-rank BatchNorm statistics stay local.

    Install with `model.enable_gradient_sync()`; `allreduce_gradients` afterwards is not needed.  Works on gloo
    (CPU tensors, no streams) for the world-size-2 tests."""

    def __init__(self, params, device, group=None):
        self.group = group
        self.params = list(params)
        self.offsets, off = {}, 0
        for p in self.params:
            self.offsets[id(p)] = (off, p.numel(), tuple(p.shape))
            off += (p.numel() + 3) // 4 * 4                      # 16-byte aligned slices
        self.flat = torch.zeros(max(off, 1), dtype=torch.float32, device=device)
        self.cuda = self.flat.is_cuda
        self.comm_stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.works, self.launched = [], set()
        self.pending, self.pending_elems = [], 0
        # Bucket size (elements).  Round 2, 8 GPUs (tools/train_bucket_probe.py, profiles/r02_train_bucket_probe_8gpu.jsonl):
        # an NCCL all-reduce over NVSwitch costs ~0.1 ms whether it carries 0.8 M or 3.3 M elements (latency-bound, NVLS),
        # and every collective is a point where the ranks wait for the slowest one -- four 1 M-element buckets took
        # 0.54-0.66 ms of collective time with 0.15-0.31 ms exposed, two buckets 0.28 / 0.12-0.13 ms, ONE bucket at the end
        # of the backward 0.10-0.13 / 0.02-0.11 ms (step 8.42-8.43 ms against 8.44-8.64 ms).  So buckets are large: the
        # 3.3 M gradients of the GraphSAGE models go in one collective, a model with more than 8 M trainable elements
        # (EA-GNN) still overlaps its first buckets with the backward.  BUCKGNN_GRAD_BUCKET_ELEMS overrides.
        self.min_bucket_elems = int(os.environ.get("BUCKGNN_GRAD_BUCKET_ELEMS", 1 << 23))
        self.events = []          # (start, end) CUDA events of every collective of the last step (timing)
        self.time_collectives = False

    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def view(self, param):
        o = self.offsets.get(id(param))
        if o is None:
            return None
        off, n, shape = o
        return self.flat[off:off + n].view(shape)

    def begin_step(self) -> None:
        self.works, self.launched, self.events = [], set(), []
        self.pending, self.pending_elems = [], 0

    def launch(self, params, force: bool = False) -> None:
        """The gradients of `params` are final: queue their slices; once enough has gathered (or at `finish`) all-reduce
        them, contiguous runs merged into one collective."""
        for p in params:
            o = self.offsets.get(id(p))
            if o is None or id(p) in self.launched:
                continue
            self.launched.add(id(p))
            self.pending.append((o[0], o[0] + (o[1] + 3) // 4 * 4))
            self.pending_elems += o[1]
        if not self.pending or (self.pending_elems < self.min_bucket_elems and not force):
            return
        spans, self.pending, self.pending_elems = self.pending, [], 0
        if self.world() == 1:
            return
        spans.sort()
        merged = [list(spans[0])]
        for a, b in spans[1:]:
            if a == merged[-1][1]:
                merged[-1][1] = b
            else:
                merged.append([a, b])
        avg = dist.ReduceOp.AVG if self.cuda else dist.ReduceOp.SUM      # gloo has no AVG
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ready)
                for a, b in merged:
                    if self.time_collectives:
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                    self.works.append(dist.all_reduce(self.flat[a:b], op=avg, group=self.group, async_op=True))
                    if self.time_collectives:
                        self.works[-1].wait()                         # comm stream waits (not the host, not compute)
                        e1.record()
                        self.events.append((e0, e1))
        else:
            for a, b in merged:
                dist.all_reduce(self.flat[a:b], op=avg, group=self.group)
                self.flat[a:b] /= self.world()

    def finish(self) -> None:
        """End of the backward: everything not launched yet goes now, then the compute stream waits for all of it."""
        self.launch(self.params, force=True)
        if self.cuda:
            for w in self.works:
                w.wait()                                              # current (compute) stream waits for the collective
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.works = []
