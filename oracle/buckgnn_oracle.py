"""Pure-torch CPU restatement of `BuckGNN.forward` (reference `Models/BuckGNN.py:311-526`).

TEST INFRASTRUCTURE ONLY -- see `oracle/__init__.py`.  Pinned: `tests/test_reference_source.py`
runs the reference's own `Models/BuckGNN.py` (unmodified, `oracle/reference_source.py`) against
this module to 1e-6.  PyG and torch_scatter are not installable here, so THEIR operators stay
restated from their published semantics at the reference's call sites (below).

Operator semantics restated
---------------------------
* `sage_conv`  <- PyG `SAGEConv(in, out, normalize=True, aggr=...)`, constructed at
  `Models/BuckGNN.py:114-176`, applied at `:342,393,434,449,463`:
  messages flow edge_index[0] (source j) -> edge_index[1] (target i);
  mean: sum / max(count, 1); add/sum: plain sum; max: element-wise max, 0 for
  nodes without in-edges; `out = lin_l(agg) + lin_r(x)` (bias only in lin_l);
  `F.normalize(out, p=2, dim=-1)` (eps 1e-12).
* `global_mean_pool` <- PyG, used at `Models/BuckGNN.py:274-291,579`:
  per-graph sum / max(count, 1), G = batch.max()+1; `batch=None` -> mean over all
  nodes, keepdim.
* `scatter_mean` <- torch_scatter, used at `Models/BuckGNN.py:561`:
  sum / max(count, 1) with `dim_size` rows.
* `OracleGraphNetBlock` <- `Models/BuckGNN.py:528-566` (aggregates on
  `row = edge_index[0]`).
* `OracleBuckGNN` <- `Models/BuckGNN.py:9-526`: same constructor, same parameter
  names/shapes/registration order (so a state_dict moves between the two), same
  forward for the model_name / pooling_layer values that work in the reference,
  including the SAGPooling variants `GraphSAGE_SAG` (`:190-217`, `:493-511`) and
  `EAGNN_SAG` (`:219-244`, `:354-373`).
* `OracleSAGPooling`, `topk`, `filter_adj` <- PyG `SAGPooling(hidden, ratio=0.5,
  GNN=SAGEConv, aggr='add')` constructed at `Models/BuckGNN.py:203-208,231-236`,
  applied at `:365-367,502-504` (PyG < 2.4 formulation, whose state_dict is
  `pool.gnn.{lin_l.weight,lin_l.bias,lin_r.weight}`):
  `score = tanh(gnn(x, edge_index).view(-1))`; per graph keep the
  `ceil(ratio * n_g)` highest scores in descending order (PyG sorts with an
  unstable sort, so ties are unspecified there; here ties keep the lower node
  id first); `x = x[perm] * score[perm]`; `batch = batch[perm]`; edges with both
  endpoints kept survive in their original order with relabelled endpoints.

`dtype=torch.float64` copies of the module give the error-budget reference.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------- operators
def scatter_sum(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)


def scatter_mean(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """torch_scatter.scatter_mean(src, index, dim=0, dim_size=dim_size)."""
    s = scatter_sum(src, index, dim_size)
    cnt = torch.bincount(index, minlength=dim_size).clamp(min=1).to(src.dtype)
    return s / cnt.view(-1, *([1] * (src.dim() - 1)))


def scatter_max(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """PyG 'max' aggregation: rows nobody scatters to stay 0."""
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return out.scatter_reduce_(0, idx, src, reduce="amax", include_self=False)


def aggregate(x: torch.Tensor, edge_index: torch.Tensor, aggr: str) -> torch.Tensor:
    src, dst = edge_index[0], edge_index[1]
    msgs = x.index_select(0, src)
    n = x.shape[0]
    if aggr == "mean":
        return scatter_mean(msgs, dst, n)
    if aggr in ("add", "sum"):
        return scatter_sum(msgs, dst, n)
    if aggr == "max":
        return scatter_max(msgs, dst, n)
    raise ValueError(f"unknown aggr {aggr!r}")


def global_mean_pool(x: torch.Tensor, batch) -> torch.Tensor:
    if batch is None:
        return x.mean(dim=0, keepdim=True)
    g = int(batch.max()) + 1
    return scatter_mean(x, batch, g)


class OracleSAGEConv(nn.Module):
    """PyG SAGEConv(in, out, normalize=..., aggr=...), root_weight=True, bias=True."""

    def __init__(self, in_channels: int, out_channels: int, normalize: bool = True, aggr: str = "mean"):
        super().__init__()
        self.aggr = aggr
        self.normalize = normalize
        self.lin_l = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x, edge_index):
        out = self.lin_l(aggregate(x, edge_index, self.aggr)) + self.lin_r(x)
        if self.normalize:
            out = F.normalize(out, p=2.0, dim=-1)
        return out


def _mlp2(i, h, o):
    return nn.Sequential(nn.Linear(i, h), nn.ReLU(), nn.Linear(h, o))


class OracleGraphNetBlock(nn.Module):
    """Reference `GraphNetBlock` (`Models/BuckGNN.py:528-566`)."""

    def __init__(self, hidden_channels: int):
        super().__init__()
        h = hidden_channels
        self.edge_mlp = _mlp2(3 * h, h, h)
        self.node_mlp_phi = _mlp2(2 * h, h, h)
        self.node_mlp_gamma = _mlp2(2 * h, h, h)
        self.node_mlp_beta = _mlp2(h, h, h)

    def forward(self, x, edge_index, edge_attr):
        row, col = edge_index[0], edge_index[1]
        edge_attr = self.edge_mlp(torch.cat([x[row], x[col], edge_attr], dim=1))
        messages = self.node_mlp_phi(torch.cat([x[col], edge_attr], dim=1))
        agg = scatter_mean(messages, row, x.size(0))
        x = self.node_mlp_gamma(torch.cat([x, agg], dim=1))
        x = x + self.node_mlp_beta(x)
        return x, edge_attr



def topk(score: torch.Tensor, ratio: float, batch: torch.Tensor) -> torch.Tensor:
    """PyG `topk(x, ratio, batch)` (torch_geometric.nn.pool.topk_pool / select.topk) for a float ratio:
    per graph the `ceil(ratio * n_g)` highest-scoring nodes, graphs in order, descending score inside a
    graph; ties resolved towards the lower node id (stable)."""
    n = score.shape[0]
    g = int(batch.max()) + 1 if n > 0 else 0
    num_nodes = torch.bincount(batch, minlength=g)
    k = (ratio * num_nodes.to(torch.float)).ceil().to(torch.long)          # as PyG computes it
    k = torch.minimum(k, num_nodes)
    order = torch.sort(score, descending=True, stable=True).indices
    order = order[torch.sort(batch[order], stable=True).indices]            # grouped by graph, descending inside
    start = torch.cumsum(num_nodes, 0) - num_nodes
    pos = torch.arange(n) - start[batch[order]]
    return order[pos < k[batch[order]]]


def filter_adj(edge_index: torch.Tensor, edge_attr, perm: torch.Tensor, num_nodes: int):
    """PyG `filter_adj`: keep the edges whose two endpoints are in `perm`, relabel them to positions in `perm`."""
    mask = perm.new_full((num_nodes,), -1)
    mask[perm] = torch.arange(perm.size(0), dtype=perm.dtype)
    row, col = mask[edge_index[0]], mask[edge_index[1]]
    keep = (row >= 0) & (col >= 0)
    if edge_attr is not None:
        edge_attr = edge_attr[keep]
    return torch.stack([row[keep], col[keep]], dim=0), edge_attr


class OracleSAGPooling(nn.Module):
    """PyG `SAGPooling(in_channels, ratio, GNN=SAGEConv, aggr=...)`: min_score=None, multiplier=1,
    nonlinearity=tanh; the scoring GNN is `SAGEConv(in_channels, 1, aggr=aggr)` (normalize=False)."""

    def __init__(self, in_channels: int, ratio: float = 0.5, aggr: str = "add"):
        super().__init__()
        self.ratio = ratio
        self.gnn = OracleSAGEConv(in_channels, 1, normalize=False, aggr=aggr)

    def forward(self, x, edge_index, edge_attr=None, batch=None):
        if batch is None:
            batch = edge_index.new_zeros(x.size(0))
        score = torch.tanh(self.gnn(x, edge_index).view(-1))
        perm = topk(score, self.ratio, batch)
        x = x[perm] * score[perm].view(-1, 1)
        batch = batch[perm]
        edge_index, edge_attr = filter_adj(edge_index, edge_attr, perm, score.size(0))
        return x, edge_index, edge_attr, batch, perm, score[perm]


class OracleMLPPooling(nn.Module):
    """Reference `MLPPooling` (`Models/BuckGNN.py:568-581`)."""

    def __init__(self, in_channels, hidden_channels, out_channels):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(in_channels, hidden_channels), nn.ReLU())

    def forward(self, x, batch):
        return self.mlp(global_mean_pool(x, batch))


_SAGE_LISTS = {  # model_name -> (ModuleList attribute, aggr)   Models/BuckGNN.py:120-180
    "GraphSage_sumAggr": ("sage_blocks_sum", "sum"),
    "GraphSage_addAggr": ("sage_blocks_add", "add"),
    "GraphSage_meanAggr": ("sage_blocks_mean", "mean"),
    "GraphSage_maxAggr": ("sage_blocks_max", "max"),
}


def output_dim_for(prediction_type, use_z_coord, use_rotations):
    """`Models/BuckGNN.py:19-38`."""
    if prediction_type == "buckling":
        return 1
    if prediction_type == "static_disp":
        return {(True, True): 6, (True, False): 3, (False, True): 4, (False, False): 2}[
            (bool(use_z_coord), bool(use_rotations))]
    if prediction_type == "static_stress":
        return 3
    if prediction_type == "mode_shape":
        return 6 if use_rotations else 3
    return None  # the reference leaves output_dim unbound -> NameError at decoder build


class OracleBuckGNN(nn.Module):
    def __init__(self, num_node_features, num_edge_features, hidden_channels=128,
                 num_layers=6, pooling_layer="mean", prediction_type="buckling",
                 use_z_coord=False, use_rotations=False, dropout_rate=0.1,
                 model_name="GraphSAGE_MLP"):
        super().__init__()
        self.hidden_channels = hidden_channels
        self.prediction_type = prediction_type
        self.pooling_layer = pooling_layer
        self.num_layers = num_layers
        self.model_name = model_name
        output_dim = output_dim_for(prediction_type, use_z_coord, use_rotations)
        h = hidden_channels
        cat_dec = pooling_layer == "supernode_with_pooling" and prediction_type == "buckling"
        if h <= 128:
            self.node_encoder = _mlp2(num_node_features, 64, h)
            self.edge_encoder = _mlp2(num_edge_features, 64, h)
            self.decoder = _mlp2(2 * h if cat_dec else h, 64, output_dim)
        elif h >= 256:
            self.node_encoder = nn.Sequential(nn.Linear(num_node_features, 64), nn.ReLU(),
                                              nn.Linear(64, 128), nn.ReLU(), nn.Linear(128, h))
            self.edge_encoder = nn.Sequential(nn.Linear(num_edge_features, 64), nn.ReLU(),
                                              nn.Linear(64, 128), nn.ReLU(), nn.Linear(128, h))
            self.decoder = nn.Sequential(nn.Linear(2 * h if cat_dec else h, 128), nn.ReLU(),
                                         nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, output_dim))
        if model_name == "EA_GNN_Shared":
            self.shared_gn_block = OracleGraphNetBlock(h)
        if model_name == "EA_GNN":
            self.gn_blocks = nn.ModuleList([OracleGraphNetBlock(h) for _ in range(num_layers)])
        if model_name == "GraphSage_addAggr_Shared":
            self.shared_graphsage_block = OracleSAGEConv(h, h, normalize=True, aggr="add")
        if model_name in _SAGE_LISTS:
            attr, aggr = _SAGE_LISTS[model_name]
            blocks, bns, mlps = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
            setattr(self, attr, blocks)
            self.batch_norms = bns
            self.sage_mlps = mlps
            for _ in range(num_layers):
                blocks.append(OracleSAGEConv(h, h, normalize=True, aggr=aggr))
                bns.append(nn.BatchNorm1d(h))
                mlps.append(nn.Linear(h, h))
        self.batch_norm = nn.BatchNorm1d(h)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(p=dropout_rate)
        self.pooling_mpl = OracleMLPPooling(h, h, h)
        if model_name in ("GraphSAGE_SAG", "EAGNN_SAG"):           # Models/BuckGNN.py:190-244
            n_before = num_layers // 2
            n_after = num_layers - n_before
            sage = model_name == "GraphSAGE_SAG"
            mk = (lambda: OracleSAGEConv(h, h, normalize=True, aggr="add")) if sage else (lambda: OracleGraphNetBlock(h))
            first = nn.ModuleList([mk() for _ in range(n_before)])
            setattr(self, "sage_layers_1" if sage else "gnn_layers_1", first)
            self.batch_norms_1 = nn.ModuleList([nn.BatchNorm1d(h) for _ in range(n_before)] if sage else [])
            self.pool = OracleSAGPooling(h, ratio=0.5, aggr="add")
            second = nn.ModuleList([mk() for _ in range(n_after)])
            setattr(self, "sage_layers_2" if sage else "gnn_layers_2", second)
            self.batch_norms_2 = nn.ModuleList([nn.BatchNorm1d(h) for _ in range(n_after)] if sage else [])

    # -- pooling, `Models/BuckGNN.py:246-307`
    def get_pooling_layer(self, x, edge_index, batch):
        p = self.pooling_layer
        if "super" in p:
            if batch is None:
                super_idx = torch.tensor([x.size(0) - 1])
            else:
                # last node of each graph is its super node (:256-266)
                change = torch.nonzero(batch[1:] != batch[:-1]).flatten()
                super_idx = torch.cat([change, torch.tensor([x.size(0) - 1])])
            real = torch.ones(x.size(0), dtype=torch.bool)
            real[super_idx] = False
            real_nodes = torch.where(real)[0]
            rb = None if batch is None else batch[real_nodes]
        if p == "mean":
            return global_mean_pool(x, batch)
        if p == "mean_no_super":
            return global_mean_pool(x[real_nodes], rb)
        if p == "supernode_only":
            return x[super_idx] if batch is not None else x[super_idx[0]]
        if p == "supernode_with_pooling":
            pooled = global_mean_pool(x[real_nodes], rb)
            return torch.cat([pooled, x[super_idx]], dim=1)
        if p == "mlp":
            return self.pooling_mpl(x, batch)
        if p == "mlp_no_super":
            return self.pooling_mpl(x[real_nodes], rb)
        raise ValueError(f"Unknown pooling layer: {p}")

    def forward(self, x, edge_index, edge_attr, batch=None, mask=None):
        name, L = self.model_name, self.num_layers
        if "super" in self.pooling_layer:
            is_real_node = x[:, -1] == 0
            real_node_batch = batch[is_real_node] if batch is not None else None
        x = self.node_encoder(x)
        if name in ("EA_GNN", "EA_GNN_Shared"):
            e = self.edge_encoder(edge_attr)
            for i in range(L):
                blk = self.shared_gn_block if name == "EA_GNN_Shared" else self.gn_blocks[i]
                x_prev, e_prev = x, e
                x, e = blk(x, edge_index, e)
                if 0 < i < L - 1:
                    x = x + x_prev
                    e = e + e_prev
                x = self.dropout(x)
                e = self.dropout(e)
        elif name == "GraphSage_addAggr_Shared":
            for i in range(L):
                x_prev = x
                x = self.relu(self.shared_graphsage_block(x, edge_index))
                if 0 < i < L - 1:
                    x = x + x_prev
                x = self.dropout(x)
        elif name in _SAGE_LISTS:
            blocks = getattr(self, _SAGE_LISTS[name][0])
            for i, (conv, bn) in enumerate(zip(blocks, self.batch_norms)):
                x_prev = x
                x = self.relu(bn(conv(x, edge_index)))
                if 0 < i < L - 1:
                    x = x + x_prev
                x = self.dropout(x)
        elif name == "EAGNN_SAG":                                  # Models/BuckGNN.py:354-373
            e = self.edge_encoder(edge_attr)
            for i, conv in enumerate(self.gnn_layers_1):
                x_prev, e_prev = x, e
                x, e = conv(x, edge_index, e)
                x = self.dropout(x)
                e = self.dropout(e)
                if i > 0:
                    x = x + x_prev
                    e = e + e_prev
            x, edge_index, e, batch, _, _ = self.pool(x, edge_index, e, batch)
            for conv in self.gnn_layers_2:
                x_prev, e_prev = x, e
                x, e = conv(x, edge_index, e)
                x = self.dropout(x)
                e = self.dropout(e)
                x = x + x_prev
                e = e + e_prev
        elif name == "GraphSAGE_SAG":                              # Models/BuckGNN.py:493-511
            for i, (conv, bn) in enumerate(zip(self.sage_layers_1, self.batch_norms_1)):
                identity = x
                x = self.dropout(self.relu(bn(conv(x, edge_index))))
                if i > 0:
                    x = x + identity
            x, edge_index, edge_attr, batch, _, _ = self.pool(x, edge_index, edge_attr, batch)
            for conv, bn in zip(self.sage_layers_2, self.batch_norms_2):
                identity = x
                x = self.dropout(self.relu(bn(conv(x, edge_index))))
                x = x + identity
        # any other model_name: encoder -> pooling -> decoder only (reference default
        # "GraphSAGE_MLP" matches no branch, Models/BuckGNN.py:12,326-511)
        if self.prediction_type == "buckling":
            pooled = self.get_pooling_layer(x, edge_index, batch)
            return self.decoder(pooled).squeeze(), batch
        if "static" in self.prediction_type or "mode_shape" in self.prediction_type:
            if "super" in self.pooling_layer:
                return self.decoder(x[is_real_node]), real_node_batch
            return self.decoder(x), batch
        raise ValueError(f"Unknown prediction type: {self.prediction_type}")


def randomize_bn_stats(model: nn.Module, seed: int = 1, realistic: bool = False) -> None:
    """Give every BatchNorm non-trivial running stats and affine terms, so a wrong
    BN fold cannot hide behind the 0/1 defaults.

    `realistic=True` uses the statistics a trained network would have after an
    L2-normalised 512-wide layer (entries ~ 1/sqrt(512)): running_var ~ U(.5,1.5)/C,
    running_mean ~ N(0, .3/sqrt(C)).  BN then rescales by ~sqrt(C), which makes the
    prediction sensitive to the message passing (and to GEMM rounding) instead of
    being dominated by the BN constants -- the honest setting for parity tests."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, nn.BatchNorm1d):
                c = m.num_features
                if realistic:
                    m.running_mean.copy_(torch.randn(c, generator=g) * 0.3 / c ** 0.5)
                    m.running_var.copy_((torch.rand(c, generator=g) + 0.5) / c)
                else:
                    m.running_mean.copy_(torch.randn(c, generator=g) * 0.1)
                    m.running_var.copy_(torch.rand(c, generator=g) + 0.5)
                m.weight.copy_(torch.rand(c, generator=g) + 0.5)
                m.bias.copy_(torch.randn(c, generator=g) * 0.1)
