"""Parity at BASELINE.json's FULL sizes (configs[1]: 256 plates, ~1.03 M nodes, ~6.1 M directed edges).

The CPU oracle needs minutes for a batch of this size, so the checks here are the size-independent
properties the domain offers, evaluated on the device (torch only as the checker's calculator):
  * CSR build: bit-exact against torch's stable sort at full size, plus sortedness / permutation invariants;
  * aggregation: linearity in fp32 storage and the degree-weighted checksum  sum_i agg_sum(x)_i = sum_j outdeg_j x_j;
  * pooling: count-weighted means add up to the global column sums;
  * forward: graphs are independent (the two half batches give the same eigenvalues as the full batch), reordering
    the graphs permutes the predictions, repeated runs are bit-identical, and a 4-graph sample of the SAME batch
    matches the fp32 oracle within rtol 1e-3;
  * SAGPooling: ceil(n/2) nodes per graph, scores non-increasing inside a graph, perm injective, every kept edge has
    both ends kept and its count equals the mask count."""
import pytest
import torch

from buckgnn_b200 import capi, engine
from buckgnn_b200.engine import Activation, build_graph_index
from buckgnn_b200.model import BuckGNN
from buckgnn_b200.synth import collate, config_batch, make_plate_graph
from oracle.buckgnn_oracle import OracleBuckGNN, randomize_bn_stats

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = 256


@pytest.fixture(scope="module")
def graphs():
    return [make_plate_graph(i) for i in range(G)]


@pytest.fixture(scope="module")
def batch(graphs):
    b = collate(graphs)
    assert b.num_graphs == G and b.num_nodes > 900_000 and b.num_edges > 5_000_000     # configs[1] scale
    return b.to(DEV)


@pytest.fixture(scope="module")
def pair():
    torch.manual_seed(0)
    cfg = dict(num_node_features=16, num_edge_features=5, hidden_channels=512, num_layers=6,
               pooling_layer="mean", model_name="GraphSage_meanAggr")
    ref = OracleBuckGNN(**cfg).eval()
    randomize_bn_stats(ref, realistic=True)
    ours = BuckGNN(**cfg)                      # default precision (fp16 operands, fp32 accumulate)
    ours.load_state_dict(ref.state_dict())
    return ref, ours.to(DEV).eval()


@pytest.mark.parametrize("key_row", [1, 0])
def test_csr_build_full_size_bit_exact_and_sorted(batch, key_row):
    n, e = batch.num_nodes, batch.num_edges
    idx = build_graph_index(batch.edge_index, batch.batch, n, key_row=key_row)
    key, other = batch.edge_index[key_row], batch.edge_index[1 - key_row]
    order = torch.sort(key, stable=True).indices
    perm = idx.perm[:e].long()
    assert torch.equal(perm, order)                                                  # == argsort(key, stable)
    assert torch.equal(idx.col[:e].long(), other[order])
    rowptr = idx.rowptr.long()
    assert torch.equal(rowptr, torch.cat([rowptr.new_zeros(1), torch.bincount(key, minlength=n).cumsum(0)]))
    # invariants that need no second implementation
    assert int(rowptr[0]) == 0 and int(rowptr[-1]) == e and bool((rowptr[1:] >= rowptr[:-1]).all())
    assert torch.equal(torch.sort(perm).values, torch.arange(e, device=DEV))         # a permutation of the edges
    ks = key[perm]
    assert bool((ks[1:] >= ks[:-1]).all())                                           # sorted by key
    assert bool(((ks[1:] != ks[:-1]) | (perm[1:] > perm[:-1])).all())                # stable inside a key
    ptr = idx.graph_ptr.long()
    assert torch.equal(ptr, batch.ptr.to(DEV)) and idx.n_graphs == G                 # pooling offsets, bit-exact
    if key_row == 1:
        assert idx.n_big == G and idx.hub_lo is not None                             # one range hub (super node) per graph


def test_aggregation_full_size_linearity_and_checksum(batch):
    n, e = batch.num_nodes, batch.num_edges
    idx = build_graph_index(batch.edge_index, batch.batch, n)
    g = torch.Generator(device=DEV).manual_seed(1)
    xs = [Activation(n, 512, "tf32", DEV) for _ in range(3)]                         # fp32 storage
    xs[0].data.copy_(torch.randn(n, 512, generator=g, device=DEV))
    xs[1].data.copy_(torch.randn(n, 512, generator=g, device=DEV))
    a, b = 0.75, -1.5                                                                # exact in fp32: the combination is too
    xs[2].data.copy_(a * xs[0].data + b * xs[1].data)
    outs = [Activation(n, 512, "tf32", DEV) for _ in range(3)]
    for aggr in ("sum", "mean"):
        for x, o in zip(xs, outs):
            engine.aggregate(x, o, idx, aggr)
        want = a * outs[0].data + b * outs[1].data
        err = (outs[2].data - want).abs().max().item()
        scale = want.abs().max().item()
        assert err <= 2e-5 * scale, (aggr, err, scale)                               # fp32 summation-order noise only
    # checksum of checksums for the sum aggregation: every x_j is counted once per out-edge
    engine.aggregate(xs[0], outs[0], idx, "sum")
    outdeg = torch.bincount(batch.edge_index[0], minlength=n).double()
    want = (outdeg[:, None] * xs[0].data.double()).sum(0)
    got = outs[0].data.double().sum(0)
    assert ((got - want).abs().max() / want.abs().max()).item() < 1e-6
    # 16-bit storage: the same checksum within the rounding of the stored aggregate
    xh, oh = Activation(n, 512, "fp16", DEV), Activation(n, 512, "fp16", DEV)
    xh.data.copy_(xs[0].data)
    engine.aggregate(xh, oh, idx, "mean")
    deg = torch.bincount(batch.edge_index[1], minlength=n).clamp(min=1).double()
    got = (oh.data.double() * deg[:, None]).sum(0)
    want = (outdeg[:, None] * xh.data.double()).sum(0)
    assert ((got - want).abs().max() / want.abs().max()).item() < 2e-3


def test_pool_full_size_counts_times_means_are_column_sums(batch, pair):
    _, ours = pair
    n = batch.num_nodes
    idx = build_graph_index(batch.edge_index, batch.batch, n)
    x = Activation(n, 512, "tf32", DEV)
    x.data.copy_(torch.randn(n, 512, generator=torch.Generator(device=DEV).manual_seed(2), device=DEV))
    _, pooled = engine.pool_head(x, idx, ours._packed()["dec"], 1, want_pooled=True)
    counts = (batch.ptr[1:] - batch.ptr[:-1]).to(DEV).double()
    got = (pooled.double() * counts[:, None]).sum(0)
    want = x.data.double().sum(0)
    assert ((got - want).abs().max() / want.abs().max()).item() < 1e-5


def test_forward_full_size_properties(graphs, batch, pair):
    ref, ours = pair
    with torch.no_grad():
        full, bb = ours(batch.x, batch.edge_index, batch.edge_attr, batch.batch)
        again, _ = ours(batch.x, batch.edge_index, batch.edge_attr, batch.batch)
    assert bb is batch.batch and full.shape == (G,) and bool(torch.isfinite(full).all())
    assert torch.equal(full, again)                                                  # deterministic kernels
    # graphs are independent: sharding the batch (what the multi-GPU path does) changes nothing but summation order
    halves = []
    for part in (graphs[:G // 2], graphs[G // 2:]):
        hb = collate(part).to(DEV)
        with torch.no_grad():
            halves.append(ours(hb.x, hb.edge_index, hb.edge_attr, hb.batch)[0])
    sharded = torch.cat(halves)
    assert ((sharded - full).abs() / full.abs().clamp(min=1e-3)).max().item() < 2e-4
    # relabelling the graphs permutes the predictions
    order = torch.randperm(G, generator=torch.Generator().manual_seed(3)).tolist()
    pb = collate([graphs[i] for i in order]).to(DEV)
    with torch.no_grad():
        permuted = ours(pb.x, pb.edge_index, pb.edge_attr, pb.batch)[0]
    assert ((permuted - full[order]).abs() / full[order].abs().clamp(min=1e-3)).max().item() < 2e-4
    # a sample of the same batch against the fp32 oracle (BASELINE.json: rtol 1e-3 on the eigenvalues)
    sample = [0, 85, 170, 255]
    sb = collate([graphs[i] for i in sample])
    with torch.no_grad():
        want, _ = ref(sb.x, sb.edge_index, sb.edge_attr, sb.batch)
    got = full[sample].cpu()
    assert ((got - want).abs() / want.abs().clamp(min=1e-3)).max().item() < 1e-3


def _sample_against_oracle(model_name, precision, graphs, sample, rtol, layers=6):
    """full batch on the device; the oracle on a few graphs of the SAME batch (graphs are independent)"""
    torch.manual_seed(1)
    cfg = dict(num_node_features=16, num_edge_features=5, hidden_channels=512, num_layers=layers,
               pooling_layer="mean", model_name=model_name)
    ref = OracleBuckGNN(**cfg).eval()
    randomize_bn_stats(ref, realistic=True)
    ours = BuckGNN(**cfg, precision=precision)
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(DEV).eval()
    b = collate(graphs).to(DEV)
    with torch.no_grad():
        full, _ = ours(b.x, b.edge_index, b.edge_attr, b.batch)
        sb = collate([graphs[i] for i in sample])
        want, _ = ref(sb.x, sb.edge_index, sb.edge_attr, sb.batch)
    assert full.shape == (len(graphs),) and bool(torch.isfinite(full).all())
    got = full[sample].cpu()
    rel = ((got - want).abs() / want.abs().clamp(min=1e-3)).max().item()
    assert rel < rtol, f"{model_name} {precision}: max rel err {rel:.3e} >= {rtol}"


def test_eagnn_configs2_size_sample_matches_oracle():
    """BASELINE.json configs[2]: EA_GNN 6x512 on 128 stiffened plates with virtual edges (~0.5 M nodes, ~5.6 M edges);
    two graphs of the batch against the fp32 oracle at rtol 1e-3 in the mode tools/bench_configs.py cfg3 reports."""
    graphs = [make_plate_graph(i, stiffened=True) for i in range(128)]
    assert sum(g.num_nodes for g in graphs) > 400_000
    _sample_against_oracle("EA_GNN", "fp16", graphs, [0, 127], 1e-3)


def test_meanaggr_configs4_mixed_mesh_scales_sample_matches_oracle():
    """BASELINE.json configs[4]'s model and data: GraphSage_meanAggr on stiffened plates with virtual edges, mesh node
    counts x1 .. x8 mixed in one batch (hub rows of 3 k .. 46 k entries; non-range hubs take the generic hub kernel)."""
    scales = [1.0, 2 ** 0.5, 2.0, 2 ** 1.5]
    graphs = [make_plate_graph(i, stiffened=True, scale=scales[i % 4]) for i in range(24)]
    _sample_against_oracle("GraphSage_meanAggr", "fp16", graphs, [0, 1, 2], 1e-3)


def test_sag_pool_full_size_invariants(batch):
    n, e = batch.num_nodes, batch.num_edges
    idx = build_graph_index(batch.edge_index, batch.batch, n)
    x = Activation(n, 512, "fp16", DEV)
    x.data.copy_(torch.randn(n, 512, generator=torch.Generator(device=DEV).manual_seed(4), device=DEV) * 0.5)
    g = torch.Generator().manual_seed(5)
    pack = {"w_l": (torch.randn(512, generator=g) / 512 ** 0.5).to(DEV), "w_r": (torch.randn(512, generator=g) / 512 ** 0.5).to(DEV),
            "bias": 0.05, "ratio": 0.5}
    res = engine.sag_pool(x, idx, idx.graph_ptr, idx.n_graphs, batch.edge_index, pack, want_kept_edges=True)
    sizes = (batch.ptr[1:] - batch.ptr[:-1]).to(DEV)
    keep = (sizes + 1) // 2
    assert res.n_nodes == int(keep.sum())
    assert torch.equal(torch.bincount(res.batch, minlength=G), keep)                 # ceil(n_g / 2) per graph
    assert bool((res.batch[1:] >= res.batch[:-1]).all())                             # pooled batch stays sorted
    same_graph = res.batch[1:] == res.batch[:-1]
    assert bool(((res.score[1:] <= res.score[:-1]) | ~same_graph).all())             # descending inside a graph
    perm = res.perm.long()
    assert torch.unique(perm).numel() == perm.numel()                                # injective
    assert torch.equal(batch.batch[perm], res.batch)                                 # batch' = batch[perm]
    assert torch.equal(res.all_scores[perm], res.score)
    new_id = res.new_id.long()
    assert torch.equal(new_id[perm], torch.arange(perm.numel(), device=DEV)) and int((new_id >= 0).sum()) == perm.numel()
    # the smallest kept score of a graph is not below its largest dropped score
    dropped = new_id < 0
    big = torch.full((G,), -2.0, device=DEV).scatter_reduce(0, batch.batch[dropped], res.all_scores[dropped], "amax")
    small = torch.full((G,), 2.0, device=DEV).scatter_reduce(0, res.batch, res.score, "amin")
    assert bool((small >= big).all())
    # filter_adj: mask count, order and relabelling
    src, dst = batch.edge_index[0], batch.edge_index[1]
    mask = (new_id[src] >= 0) & (new_id[dst] >= 0)
    assert res.n_edges == int(mask.sum())
    assert torch.equal(res.kept_edge[:res.n_edges].long(), torch.nonzero(mask).flatten())
    assert torch.equal(res.edge_index, torch.stack([new_id[src[mask]], new_id[dst[mask]]]))
