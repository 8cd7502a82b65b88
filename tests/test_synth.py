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
