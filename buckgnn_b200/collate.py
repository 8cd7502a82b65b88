"""Dataset-resident batching.

The reference builds every batch on the host (PyG `DataLoader` -> `Batch.from_data_list`) and then copies it
to the GPU (`batch.to(device)`, INFERENCE.py:135, TRAIN_FINAL.py:255): 294 MB over PCIe per 256-graph batch.
A B200 has 180 GB of HBM -- the whole 80 k-graph inference set of BASELINE.json configs[4] (~90 GB) fits --
so `DeviceGraphStore` keeps the graphs on the device in concatenated form and `batch(indices)` assembles a
PyG-layout batch with two kernels (`bg_collate_ptr`, `bg_collate`): no host collate, no per-batch H2D.
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import capi
from .engine import _stream
from .synth import PlateBatch, PlateGraph


class DeviceGraphStore:
    def __init__(self, graphs: Sequence[PlateGraph], device):
        dev = torch.device(device)
        n = torch.tensor([g.num_nodes for g in graphs], dtype=torch.int64)
        e = torch.tensor([g.num_edges for g in graphs], dtype=torch.int64)
        node_ptr = torch.zeros(len(graphs) + 1, dtype=torch.int64)
        edge_ptr = torch.zeros(len(graphs) + 1, dtype=torch.int64)
        node_ptr[1:] = torch.cumsum(n, 0)
        edge_ptr[1:] = torch.cumsum(e, 0)
        self.num_graphs = len(graphs)
        self.x = torch.cat([g.x for g in graphs], 0).to(torch.float32).contiguous().to(dev)
        self.edge_index = torch.cat([g.edge_index for g in graphs], 1).to(torch.int64).contiguous().to(dev)   # local ids
        self.edge_attr = torch.cat([g.edge_attr for g in graphs], 0).to(torch.float32).contiguous().to(dev)
        self.y = torch.cat([g.y.reshape(-1)[:1] for g in graphs], 0).to(torch.float32).contiguous().to(dev)
        self.node_ptr, self.edge_ptr = node_ptr.to(dev), edge_ptr.to(dev)
        self._n_host, self._e_host = n, e            # per-graph sizes on the host: a CPU `indices` needs no read-back
        self.device = dev

    def bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.x, self.edge_index, self.edge_attr, self.y))

    def batch(self, indices: torch.Tensor) -> PlateBatch:
        """PyG-layout batch of the graphs `indices` (int64), in that order.  With `indices` on the HOST the batch's
        node / edge totals come from the host-side size table and nothing is read back from the device (the loop
        never waits for the previous forward); with device indices one 16-byte read-back supplies them."""
        dev = self.device
        host_idx = indices if not indices.is_cuda else None
        sel = indices.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
        g = sel.numel()
        s = _stream()
        i64 = dict(dtype=torch.int64, device=dev)
        out_np, out_ep = torch.empty(g + 1, **i64), torch.empty(g + 1, **i64)
        capi.collate_ptr(sel.data_ptr(), g, self.node_ptr.data_ptr(), self.edge_ptr.data_ptr(), out_np.data_ptr(),
                         out_ep.data_ptr(), s)
        if host_idx is not None:
            hi = host_idx.to(torch.int64)
            n_out, e_out = int(self._n_host[hi].sum()), int(self._e_host[hi].sum())
        else:
            n_out, e_out = (int(v) for v in torch.stack([out_np[g], out_ep[g]]).tolist())     # one small read-back
        f, fe = self.x.shape[1], self.edge_attr.shape[1]
        x = torch.empty((n_out, f), dtype=torch.float32, device=dev)
        ei = torch.empty((2, e_out), **i64)
        ea = torch.empty((e_out, fe), dtype=torch.float32, device=dev)
        batch = torch.empty(n_out, **i64)
        y = torch.empty(g, dtype=torch.float32, device=dev)
        capi.collate(self.x.data_ptr(), f, self.edge_index.data_ptr(), self.edge_index.shape[1], self.edge_attr.data_ptr(), fe,
                     self.y.data_ptr(), sel.data_ptr(), g, self.node_ptr.data_ptr(), self.edge_ptr.data_ptr(),
                     out_np.data_ptr(), out_ep.data_ptr(), n_out, e_out, x.data_ptr(), ei.data_ptr(), ea.data_ptr(),
                     batch.data_ptr(), y.data_ptr(), s)
        return PlateBatch(x, ei, ea, batch, y, out_np, g)
