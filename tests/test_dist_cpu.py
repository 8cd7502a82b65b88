"""Graph sharding across ranks: partition logic + the world_size-2 gather path on gloo (CPU).
The per-shard `forward` is a stand-in (per-graph feature sum) -- the CUDA forward itself is
covered by the -m gpu tests; what is tested here is that sharding + gather reproduce the
single-process result in the original graph order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from buckgnn_b200.dist import allreduce_gradients, graph_cost, partition_graphs, sharded_predict
from buckgnn_b200.synth import make_plate_graph


def test_partition_is_balanced_deterministic_and_complete():
    costs = [graph_cost(n, 6 * n) for n in (4096, 8192, 16384, 32768, 4096, 4096, 8192, 4096, 30000, 5000)]
    for world in (1, 2, 4, 8):
        parts = partition_graphs(costs, world)
        assert sorted(i for p in parts for i in p) == list(range(len(costs)))
        assert parts == partition_graphs(costs, world)
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs)          # LPT bound
        assert all(p == sorted(p) for p in parts)


def _fake_forward(batch):
    s = torch.zeros(batch.num_graphs, dtype=torch.float32)
    return s.index_add_(0, batch.batch, batch.x.sum(1))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    graphs = [make_plate_graph(i, nx=4 + (i % 5), ny=3 + (i % 3)) for i in range(7)]
    out = sharded_predict(graphs, _fake_forward, "cpu")
    q.put((rank, out.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_process():
    graphs = [make_plate_graph(i, nx=4 + (i % 5), ny=3 + (i % 3)) for i in range(7)]
    want = [float(g.x.sum()) for g in graphs]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for r in (0, 1):
        assert torch.allclose(torch.tensor(got[r]), torch.tensor(want), rtol=1e-6)


def test_single_process_path():
    graphs = [make_plate_graph(i, nx=4, ny=4) for i in range(3)]
    out = sharded_predict(graphs, _fake_forward, "cpu")
    assert torch.allclose(out, torch.tensor([float(g.x.sum()) for g in graphs]), rtol=1e-6)


def _grad_worker(rank, world, port, q):
    """Data-parallel step on the CPU oracle: each rank backpropagates its shard, gradients are
    all-reduced (mean) and must equal the single-process gradient of the mean loss over all graphs
    (equal shard sizes, BN in eval mode so the statistics do not differ between the two runs)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.buckgnn_oracle import OracleBuckGNN
    from buckgnn_b200.synth import collate
    torch.manual_seed(0)
    model = OracleBuckGNN(16, 5, 256, 2, "mean", model_name="GraphSage_meanAggr", dropout_rate=0.0).double().eval()
    graphs = [make_plate_graph(i, nx=4, ny=3) for i in range(4)]
    mine = collate(graphs[rank * 2:rank * 2 + 2])
    pred, _ = model(mine.x.double(), mine.edge_index, mine.edge_attr.double(), mine.batch)
    ((pred - mine.y.double()) ** 2).mean().backward()
    n = allreduce_gradients(model.parameters())
    unused = [k for k, p in model.named_parameters() if p.grad is None]
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
    q.put((rank, n, sorted(set(k.split(".")[0] for k in unused)), flat.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_world_size_2_matches_single_process():
    from oracle.buckgnn_oracle import OracleBuckGNN
    from buckgnn_b200.synth import collate
    torch.manual_seed(0)
    model = OracleBuckGNN(16, 5, 256, 2, "mean", model_name="GraphSage_meanAggr", dropout_rate=0.0).double().eval()
    graphs = [make_plate_graph(i, nx=4, ny=3) for i in range(4)]
    b = collate(graphs)
    pred, _ = model(b.x.double(), b.edge_index, b.edge_attr.double(), b.batch)
    ((pred - b.y.double()) ** 2).mean().backward()
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=180) for _ in range(2)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, n, unused, flat in got:
        assert n == want.numel()
        assert unused == ["batch_norm", "edge_encoder", "pooling_mpl", "sage_mlps"]   # never touched by this model_name
        assert torch.allclose(torch.tensor(flat, dtype=torch.float64), want, rtol=1e-9, atol=1e-12)


def _sync_worker(rank, world, port, q):
    """dist.GradSync on gloo: gradients written into the flat buffer group by group, every group all-reduced (mean) as
    soon as it is declared done, shared parameters only once, the result handed out as per-step snapshots."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from buckgnn_b200.dist import GradSync
    from buckgnn_b200.train import GradStore
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in ((3, 5), (7,), (4, 4), (2, 3))]
    sync = GradSync(params, "cpu")
    out = []
    for step in range(2):
        sync.min_bucket_elems = 1 if step == 0 else 1 << 23       # every group its own collective / one collective at finish
        store = GradStore("cpu", sync)
        g0, acc0 = store.get(params[0])
        g0.copy_(torch.full((3, 5), float(rank + 1 + step)))
        store.done([params[0]])                                   # first group goes while the "backward" continues
        g2, _ = store.get(params[2])
        g2.copy_(torch.arange(16.).view(4, 4) * (rank + 1))
        z = store.zeros(params[3])
        z[:, 1:] += float(10 * (rank + 1))
        store.done([params[2], params[0]])                        # params[0] again: must not be reduced twice
        store.finish()                                            # params[3] was never declared done: reduced here
        res = store.for_params(params)
        assert res[1] is None and not acc0                        # params[1] got no gradient this step
        out.append([None if t is None else t.tolist() for t in res])
        assert res[0].data_ptr() != sync.view(params[0]).data_ptr()   # a snapshot, not the reused buffer
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_sync_buckets_world_size_2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert got[0] == got[1]                                       # both ranks hold the same (mean) gradients
    for step in range(2):
        g0, g1, g2, g3 = got[0][step]
        assert g1 is None
        assert torch.allclose(torch.tensor(g0), torch.full((3, 5), 1.5 + step))
        assert torch.allclose(torch.tensor(g2), torch.arange(16.).view(4, 4) * 1.5)
        want3 = torch.zeros(2, 3)
        want3[:, 1:] = 15.0
        assert torch.allclose(torch.tensor(g3), want3)


def test_batches_by_node_budget_keeps_order_and_covers_every_graph():
    from buckgnn_b200.dist import batches_by_node_budget
    nodes = [4000, 8000, 16000, 32000, 4000, 4000, 50000, 1000]
    idx = [0, 2, 3, 4, 6, 7]
    out = batches_by_node_budget(idx, nodes, 40000)
    assert [i for b in out for i in b] == idx                       # order kept, nothing lost or repeated
    assert out == [[0, 2], [3, 4], [6], [7]]                        # a graph above the budget rides alone
    assert all(sum(nodes[i] for i in b) <= 40000 or len(b) == 1 for b in out)
    assert batches_by_node_budget([], nodes, 10) == []
