"""Synthetic plate generator: layout facts the hot path relies on."""
import torch

from buckgnn_b200.synth import collate, config_batch, make_batch, make_plate_graph


def test_fixed_grid_sizes_match_survey():
    # 64x64 grid + super node: n+1 = 4097 nodes, 2*(2*64*63 + 4096) = 24320 directed edges
    g = make_plate_graph(0, nx=64, ny=64)
    assert g.x.shape == (4097, 16) and g.edge_index.shape == (2, 24320) and g.edge_attr.shape == (24320, 5)
    b = config_batch(0, fixed_grid=True)
    assert b.num_nodes == 65552 and b.num_edges == 389120 and b.num_graphs == 16


def test_directed_pairs_adjacent_and_features_shared():
    g = make_plate_graph(3, stiffened=True)
    ei, ea = g.edge_index, g.edge_attr
    assert torch.equal(ei[0, 0::2], ei[1, 1::2]) and torch.equal(ei[1, 0::2], ei[0, 1::2])
    assert torch.equal(ea[0::2], ea[1::2])
    # no duplicate undirected edges
    n = g.num_nodes
    key = torch.minimum(ei[0, 0::2], ei[1, 0::2]) * n + torch.maximum(ei[0, 0::2], ei[1, 0::2])
    assert key.unique().numel() == key.numel()


def test_super_node_is_last_and_hub():
    g = make_plate_graph(1)
    n = g.num_nodes - 1
    assert g.x[n, -1] == 1 and g.x[n, :-1].abs().sum() == 0 and g.x[:n, -1].abs().sum() == 0
    deg = torch.bincount(g.edge_index[1], minlength=n + 1)
    assert deg[n] == n                      # hub in-degree = graph size
    assert deg[:n].max() <= 5 and deg[:n].min() >= 3   # 2..4 mesh neighbours + hub


def test_deterministic_and_rank_disjoint():
    a, b = make_plate_graph(7), make_plate_graph(7)
    assert torch.equal(a.x, b.x) and torch.equal(a.edge_index, b.edge_index)
    assert not torch.equal(make_plate_graph(7).x[:50], make_plate_graph(8).x[:50])


def test_collate_matches_pyg_batch_layout():
    gs = [make_plate_graph(i, nx=5 + i, ny=4) for i in range(3)]
    b = collate(gs)
    assert b.ptr.tolist() == [0, 21, 46, 75]
    assert torch.equal(b.batch, torch.repeat_interleave(torch.arange(3), torch.tensor([21, 25, 29])))
    off = 0
    e0 = 0
    for i, g in enumerate(gs):
        assert torch.equal(b.edge_index[:, e0:e0 + g.num_edges], g.edge_index + off)
        off += g.num_nodes
        e0 += g.num_edges
    # block diagonal: no edge crosses graphs
    assert torch.equal(b.batch[b.edge_index[0]], b.batch[b.edge_index[1]])


def test_stiffened_density():
    g = make_plate_graph(2, stiffened=True, nx=64, ny=64)
    n = 64 * 64
    # ~4n mesh edges (sides + diagonals) + 13.33% virtual + n hub, directed x2  => ~11n
    assert 10.5 * n < g.num_edges < 11.5 * n
    assert (g.edge_attr[:, 0] == 1.0).sum() >= 20        # active stiffener edges (directed)


def _expand_wire_host(w):
    """What bg_expand_wire does on the device (capi.cu k_expand_wire), as a host loop: the spec of the wire format."""
    ei = torch.empty((2, w.num_edges), dtype=torch.int64)
    batch = torch.empty(w.num_nodes, dtype=torch.int64)
    for g in range(w.num_graphs):
        n0, n1 = int(w.node_ptr[g]), int(w.node_ptr[g + 1])
        w0, w1 = int(w.wire_ptr[g]), int(w.wire_ptr[g + 1])
        f0, f1 = int(w.full_ptr[g]), int(w.full_ptr[g + 1])
        ei[:, f0:f0 + (w1 - w0)] = w.edges[:, w0:w1].to(torch.int64)
        pairs = ((f1 - f0) - (w1 - w0)) // 2
        s, base = n1 - 1, f0 + (w1 - w0)
        others = torch.arange(n0, n0 + pairs)
        ei[0, base:base + 2 * pairs:2], ei[1, base:base + 2 * pairs:2] = s, others
        ei[0, base + 1:base + 2 * pairs:2], ei[1, base + 1:base + 2 * pairs:2] = others, s
        batch[n0:n1] = g
    return ei, batch


def test_wire_format_host_conversion_round_trips():
    """pipeline.WireBatch.from_batch (loader side): hub pairs leave the wire only where they are exactly the
    reference's block; an edge list that is not grouped by graph ships every edge unchanged."""
    from buckgnn_b200.pipeline import WireBatch
    from buckgnn_b200.synth import PlateBatch
    b = make_batch(4, nx=7, ny=5)
    w = WireBatch.from_batch(b)
    assert w.edges.shape[1] == b.num_edges - 2 * (b.num_nodes - b.num_graphs)
    ei, batch = _expand_wire_host(w)
    assert torch.equal(ei, b.edge_index) and torch.equal(batch, b.batch)
    order = torch.randperm(b.num_edges, generator=torch.Generator().manual_seed(3))
    mixed = PlateBatch(b.x, b.edge_index[:, order].contiguous(), b.edge_attr[order], b.batch, b.y, b.ptr, b.num_graphs)
    w = WireBatch.from_batch(mixed)
    assert w.edges.shape[1] == b.num_edges and torch.equal(w.wire_ptr, w.full_ptr)
    ei, batch = _expand_wire_host(w)
    assert torch.equal(ei, mixed.edge_index) and torch.equal(batch, b.batch)
    # no super node at all
    plain = make_batch(2, nx=6, ny=4, super_node=False)
    w = WireBatch.from_batch(plain)
    assert w.edges.shape[1] == plain.num_edges
    ei, _ = _expand_wire_host(w)
    assert torch.equal(ei, plain.edge_index)
