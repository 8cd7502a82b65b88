"""Synthetic plate-mesh graphs in the layout the reference's dataset code emits.

The reference builds its graphs from Nastran BDF/OP2 files with pyNastran
(`Dataset_Preparation/GraphCreate.py:143-432`), none of which can run here.  This
module restates only the *layout* of what that code produces, on structured quad
grids, so the hot path sees inputs of the right shape and statistics:

* nodes row-major, one super node appended last per graph
  (`GraphCreate.py:403-415`, `VirtualEdgeCreate.py:81-113`);
* undirected mesh edges = quad sides in element order, first occurrence wins
  (`GraphCreate.py:334-350`); stiffened plates add CBAR edges on every quad side
  and both diagonals (`Data_Generation/Data_Generation_v3.py:216-270`);
* optional random virtual edges, 13.33 % of the mesh-edge count
  (`VirtualEdgeCreate.py:21-49`);
* hub edges `(n, i)` for every real node i (`VirtualEdgeCreate.py:106-111`);
* every undirected edge emitted as the adjacent directed pair (a,b),(b,a) carrying
  the same feature row (`GraphCreate.py:417-422`);
* x  [n+1,16] = [x,y | spc | fx,fy | is_boundary | 4 stiffener bins | dx,dy |
  sx,sy,txy | is_super]  (`GraphCreate.py:182-289`);
* edge_attr [E,5] = [type, dist/1000, dir_x, dir_y, is_virtual]
  (`GraphCreate.py:350,366-369`, `VirtualEdgeCreate.py:55-77`).

`collate` restates the PyG `Batch.from_data_list` layout (concatenate x /
edge_attr, offset edge_index by the node prefix sum, build the sorted `batch`
vector) so no torch_geometric is needed.

Everything is deterministic in (seed, graph index).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

BASE_SEED = 20250301
NUM_NODE_FEATURES = 16
NUM_EDGE_FEATURES = 5


@dataclass
class PlateGraph:
    """One graph, PyG `Data`-like (x, edge_index, edge_attr, y)."""
    x: torch.Tensor           # [n+1, 16] float32
    edge_index: torch.Tensor  # [2, E] int64
    edge_attr: torch.Tensor   # [E, 5] float32
    y: torch.Tensor           # [1] float32

    @property
    def num_nodes(self) -> int:
        return self.x.shape[0]

    @property
    def num_edges(self) -> int:
        return self.edge_index.shape[1]


@dataclass
class PlateBatch:
    """PyG `Batch`-like container: what `DataLoader` hands to the driver scripts."""
    x: torch.Tensor
    edge_index: torch.Tensor
    edge_attr: torch.Tensor
    batch: torch.Tensor       # [N] int64, sorted
    y: torch.Tensor           # [G]
    ptr: torch.Tensor         # [G+1] int64 node offsets
    num_graphs: int

    def to(self, device, non_blocking: bool = False) -> "PlateBatch":
        f = lambda t: t.to(device, non_blocking=non_blocking)
        return PlateBatch(f(self.x), f(self.edge_index), f(self.edge_attr), f(self.batch),
                          f(self.y), f(self.ptr), self.num_graphs)

    def pin_memory(self) -> "PlateBatch":
        f = lambda t: t.pin_memory()
        return PlateBatch(f(self.x), f(self.edge_index), f(self.edge_attr), f(self.batch),
                          f(self.y), f(self.ptr), self.num_graphs)

    @property
    def num_nodes(self) -> int:
        return self.x.shape[0]

    @property
    def num_edges(self) -> int:
        return self.edge_index.shape[1]


def _first_occurrence_unique(keys: np.ndarray) -> np.ndarray:
    """Indices of the first occurrence of each key, in order of appearance
    (what inserting into a Python dict with `if edge not in edges` keeps)."""
    _, first = np.unique(keys, return_index=True)
    first.sort()
    return first


def _quad_side_edges(nx: int, ny: int, diagonals: bool) -> np.ndarray:
    """Undirected mesh edges [M,2] (a<b) in the reference's insertion order."""
    ix, iy = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), indexing="xy")
    a = (iy * nx + ix).ravel()          # elements row-major
    b = a + 1
    c = a + nx + 1
    d = a + nx
    # CQUAD4 sides in node order (a,b),(b,c),(c,d),(d,a)  GraphCreate.py:336-337
    cand = [np.stack([a, b], 1), np.stack([b, c], 1), np.stack([c, d], 1), np.stack([d, a], 1)]
    per_elem = np.stack(cand, 1)        # [Q,4,2]
    if diagonals:
        # CBARs on both diagonals, after the element's sides
        diag = np.stack([np.stack([a, c], 1), np.stack([b, d], 1)], 1)
        per_elem = np.concatenate([per_elem, diag], 1)
    e = per_elem.reshape(-1, 2)
    e = np.sort(e, axis=1)              # tuple(sorted([idx1, idx2]))
    n = nx * ny
    keep = _first_occurrence_unique(e[:, 0].astype(np.int64) * n + e[:, 1])
    return e[keep]


def make_plate_graph(index: int, *, stiffened: bool = False, super_node: bool = True,
                     virtual_edges: Optional[bool] = None, nx: Optional[int] = None,
                     ny: Optional[int] = None, scale: float = 1.0,
                     base_seed: int = BASE_SEED) -> PlateGraph:
    """Graph number `index` of the synthetic set.

    nx, ny default to U{48..80} drawn from the graph's own generator; `scale`
    multiplies both (config 5's 1x..8x node-count sweep uses sqrt steps).
    `virtual_edges` defaults to `stiffened` (the cfg 3/5 "denser virtual edges").
    """
    if virtual_edges is None:
        virtual_edges = stiffened
    rng = np.random.default_rng(base_seed + index)
    if nx is None:
        nx = int(rng.integers(48, 81))
    if ny is None:
        ny = int(rng.integers(48, 81))
    nx = max(2, int(round(nx * scale)))
    ny = max(2, int(round(ny * scale)))
    n = nx * ny

    # --- geometry: plate 700-1000 mm (Shape_Generation.py:388-390), centred frame
    lx, ly = rng.uniform(700.0, 1000.0, size=2)
    gx, gy = np.meshgrid(np.linspace(-0.5, 0.5, nx), np.linspace(-0.5, 0.5, ny), indexing="xy")
    coords = np.stack([gx.ravel() * lx, gy.ravel() * ly], 1)            # mm, row-major

    # --- node features
    x = np.zeros((n + (1 if super_node else 0), NUM_NODE_FEATURES), dtype=np.float32)
    half = max(lx, ly) / 2.0
    x[:n, 0:2] = coords / half                                            # ~U(-1,1)
    iy, ix = np.divmod(np.arange(n), nx)
    boundary = (ix == 0) | (ix == nx - 1) | (iy == 0) | (iy == ny - 1)
    # SPC: clamped edge = 1, simply supported edge = 0.25 (GraphCreate.py:190-197)
    spc = np.zeros(n, dtype=np.float32)
    spc[(ix == 0)] = 1.0 if rng.random() < 0.5 else 0.25
    spc[(iy == 0) & (spc == 0)] = 0.25
    x[:n, 2] = spc
    # forces on ~15 boundary nodes of the loaded edge
    loaded = np.flatnonzero(ix == nx - 1)
    k = min(len(loaded), int(rng.integers(10, 21)))
    start = int(rng.integers(0, len(loaded) - k + 1))
    x[loaded[start:start + k], 3:5] = np.clip(rng.normal(0, 1, size=(k, 2)), -5, 5)
    x[:n, 5] = boundary.astype(np.float32)
    # displacement + stress channels: robust-scaled in the reference -> ~N(0,1) clipped
    x[:n, 10:15] = np.clip(rng.normal(0, 1, size=(n, 5)), -5, 5).astype(np.float32)

    # --- mesh edges
    und = _quad_side_edges(nx, ny, diagonals=stiffened)                   # [M,2]
    etype = np.full(len(und), 0.01, dtype=np.float32)
    if stiffened:
        # 10-100 active stiffener edges in straight chains of 5-25 along grid lines
        n_active_target = int(rng.integers(10, 101))
        active = np.zeros(n, dtype=bool)                                  # marks chain nodes
        active_pairs = set()
        total = 0
        while total < n_active_target:
            ln = int(rng.integers(5, 26))
            if rng.random() < 0.5:                                        # along x
                ln = min(ln, nx - 1)
                jy = int(rng.integers(0, ny)); jx = int(rng.integers(0, nx - ln))
                nodes = jy * nx + jx + np.arange(ln + 1)
            else:                                                         # along y
                ln = min(ln, ny - 1)
                jx = int(rng.integers(0, nx)); jy = int(rng.integers(0, ny - ln))
                nodes = (jy + np.arange(ln + 1)) * nx + jx
            for p, q in zip(nodes[:-1], nodes[1:]):
                active_pairs.add((int(min(p, q)), int(max(p, q))))
            active[nodes] = True
            total += ln
        keys = und[:, 0].astype(np.int64) * n + und[:, 1]
        akeys = np.array([p * n + q for p, q in active_pairs], dtype=np.int64)
        etype[np.isin(keys, akeys)] = 1.0                                 # pid 900 (GraphCreate.py:366-367)
        # stiffener direction bins /3 (Transformation.py:5-76): count of active CBARs
        # at the node per 45-degree bin; on a grid only bins 0 (x) and 2 (y) are hit
        for (p, q) in active_pairs:
            b = 0 if (q - p) == 1 else 2
            x[p, 6 + b] += 1.0 / 3.0
            x[q, 6 + b] += 1.0 / 3.0
    n_mesh = len(und)

    is_virtual = np.zeros(n_mesh, dtype=np.float32)
    if virtual_edges:
        n_virt = int(n_mesh * 0.1333)
        existing = set((und[:, 0].astype(np.int64) * n + und[:, 1]).tolist())
        vlist = []
        while len(vlist) < n_virt:
            cand = rng.integers(0, n, size=(2 * (n_virt - len(vlist)) + 8, 2))
            cand = cand[cand[:, 0] != cand[:, 1]]
            cand.sort(axis=1)
            for p, q in cand:
                key = int(p) * n + int(q)
                if key not in existing:
                    existing.add(key)
                    vlist.append((int(p), int(q)))
                    if len(vlist) == n_virt:
                        break
        virt = np.array(vlist, dtype=und.dtype).reshape(-1, 2)
        und = np.concatenate([und, virt], 0)
        etype = np.concatenate([etype, np.zeros(len(virt), np.float32)])
        is_virtual = np.concatenate([is_virtual, np.ones(len(virt), np.float32)])

    pos = coords
    if super_node:
        x[n, :] = 0.0
        x[n, -1] = 1.0                                                    # node-type flag
        hub = np.stack([np.full(n, n, dtype=und.dtype), np.arange(n, dtype=und.dtype)], 1)
        und = np.concatenate([und, hub], 0)                               # (super, i): not sorted, as in the reference
        etype = np.concatenate([etype, np.zeros(n, np.float32)])
        is_virtual = np.concatenate([is_virtual, np.ones(n, np.float32)])
        pos = np.concatenate([coords, np.zeros((1, 2))], 0)               # super node at the origin

    d = pos[und[:, 1]] - pos[und[:, 0]]
    dist = np.sqrt((d ** 2).sum(1))
    dist_safe = np.where(dist > 0, dist, 1.0)
    feat = np.stack([etype, (dist / 1000.0).astype(np.float32),
                     (d[:, 0] / dist_safe).astype(np.float32),
                     (d[:, 1] / dist_safe).astype(np.float32), is_virtual], 1).astype(np.float32)

    # directed pairs (a,b),(b,a) adjacent, same feature row for both
    m = len(und)
    ei = np.empty((2, 2 * m), dtype=np.int64)
    ei[0, 0::2] = und[:, 0]; ei[1, 0::2] = und[:, 1]
    ei[0, 1::2] = und[:, 1]; ei[1, 1::2] = und[:, 0]
    ea = np.repeat(feat, 2, axis=0)

    y = np.float32(rng.normal())
    return PlateGraph(torch.from_numpy(x), torch.from_numpy(ei), torch.from_numpy(ea),
                      torch.tensor([y], dtype=torch.float32))


def collate(graphs: Sequence[PlateGraph]) -> PlateBatch:
    """PyG `Batch.from_data_list` layout without PyG."""
    sizes = [g.num_nodes for g in graphs]
    ptr = torch.zeros(len(graphs) + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(torch.tensor(sizes, dtype=torch.int64), 0)
    x = torch.cat([g.x for g in graphs], 0)
    ea = torch.cat([g.edge_attr for g in graphs], 0)
    ei = torch.cat([g.edge_index + int(ptr[i]) for i, g in enumerate(graphs)], 1)
    batch = torch.repeat_interleave(torch.arange(len(graphs), dtype=torch.int64),
                                    torch.tensor(sizes, dtype=torch.int64))
    y = torch.cat([g.y for g in graphs], 0)
    return PlateBatch(x, ei.contiguous(), ea, batch, y, ptr, len(graphs))


def make_batch(num_graphs: int, *, first_index: int = 0, **kw) -> PlateBatch:
    return collate([make_plate_graph(first_index + i, **kw) for i in range(num_graphs)])


# BASELINE.json `configs`, by position.
def config_batch(cfg: int, *, rank: int = 0, num_graphs: Optional[int] = None,
                 fixed_grid: bool = False) -> PlateBatch:
    """The synthetic batch of BASELINE.json configs[cfg] (cfg 0..3); `rank` shifts
    the graph indices so each rank of a sharded run sees different graphs."""
    g_default = {0: 16, 1: 256, 2: 128, 3: 16}[cfg]
    g = g_default if num_graphs is None else num_graphs
    kw = dict(first_index=rank * g_default)
    if fixed_grid:
        kw.update(nx=64, ny=64)
    if cfg == 2:
        kw.update(stiffened=True)
    return make_batch(g, **kw)
