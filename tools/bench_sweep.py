"""BASELINE.json configs[4]: throughput sweep over a synthetic STIFFENED inference set whose mesh sizes are scaled
1x .. 8x (node count; the super node's hub degree = graph size grows with it), graph-sharded over the ranks.

  python tools/bench_sweep.py [--graphs 80000] [--pool 256] [--node-budget 1000000]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         tools/bench_sweep.py --gpus N ...

What runs:
  * the set: `--graphs` graphs, graph j being mesh number j % pool of a pool of `--pool` distinct generated stiffened
    plates (mesh scale 1, sqrt2, 2, 2*sqrt2 by j % 4 -> node counts x1, x2, x4, x8).  Generating 80 k distinct meshes
    on the host takes minutes and ~50 GB of host memory and changes nothing for the GPU, so the pool is generated
    once and lives on every rank's device (`collate.DeviceGraphStore`, row f1);
  * sharding: `dist.partition_graphs` (greedy longest-processing-time on E + 2N) over ALL graphs of the set,
    no data-path collective; the reported imbalance is max / mean of the per-rank cost;
  * each rank cuts its shard into batches of at most `--node-budget` nodes (`dist.batches_by_node_budget`),
    assembles every batch on the device (`bg_collate`) and runs GraphSage_meanAggr 6x512 (fp16 operands) on it;
  * timing: barrier + synchronize on both sides, CUDA events, max over ranks; rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from buckgnn_b200 import capi
from buckgnn_b200.collate import DeviceGraphStore
from buckgnn_b200.dist import batches_by_node_budget, graph_cost, partition_graphs
from buckgnn_b200.synth import make_plate_graph
from tools.bench_configs import seeded_model

SCALES = (1.0, 2 ** 0.5, 2.0, 2 * 2 ** 0.5)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--graphs", type=int, default=80000)
    ap.add_argument("--pool", type=int, default=256)
    ap.add_argument("--node-budget", type=int, default=1_000_000)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="fp16")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    capi.device_check()

    pool = [make_plate_graph(i, stiffened=True, scale=SCALES[i % 4]) for i in range(args.pool)]
    store = DeviceGraphStore(pool, dev)
    nodes = [g.num_nodes for g in pool]
    edges = [g.num_edges for g in pool]
    member = [j % args.pool for j in range(args.graphs)]                    # graph j of the set -> mesh of the pool
    costs = [graph_cost(nodes[m], edges[m]) for m in member]
    parts = partition_graphs(costs, world)
    loads = [sum(costs[j] for j in p) for p in parts]
    mine = parts[rank]
    batches = batches_by_node_budget(mine, [nodes[m] for m in member], args.node_budget)
    sels = [torch.tensor([member[j] for j in b], dtype=torch.int64).pin_memory() for b in batches]      # host indices: no read-back

    cfg = dict(num_node_features=16, num_edge_features=5, hidden_channels=512, num_layers=6,
               pooling_layer="mean", model_name="GraphSage_meanAggr")
    model = seeded_model(cfg, args.precision, device=dev)

    def run(sel):
        b = store.batch(sel)
        pred, _ = model(b.x, b.edge_index, b.edge_attr, b.batch)
        return pred

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    preds = []
    with torch.no_grad():
        for sel in sels[:args.warmup]:
            run(sel)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for sel in sels:
            preds.append(run(sel))
        e1.record()
        barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    done = torch.tensor([sum(p.numel() for p in preds)], dtype=torch.int64, device=dev)
    finite = torch.tensor([int(all(bool(torch.isfinite(p).all()) for p in preds))], dtype=torch.int64, device=dev)
    ms_all = ms.clone()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(done, op=dist.ReduceOp.SUM)
        dist.all_reduce(finite, op=dist.ReduceOp.MIN)
        gathered = [torch.zeros_like(ms_all) for _ in range(world)]
        dist.all_gather(gathered, ms_all)
        per_rank_ms = [float(t.item()) for t in gathered]
    else:
        per_rank_ms = [float(ms.item())]
    # the same mesh must predict the same eigenvalue whatever batch it rode in (graphs are independent)
    first = {}
    consistent = True
    for b, p in zip(batches, preds):
        v = p.reshape(-1).float().cpu()
        for j, val in zip(b, v.tolist()):
            m = member[j]
            if m in first:
                consistent &= abs(first[m] - val) <= 1e-3 * max(abs(first[m]), 1e-6)
            else:
                first[m] = val
    if rank == 0:
        sec = float(ms.item()) * 1e-3
        tot_nodes = sum(nodes[m] for m in member)
        tot_edges = sum(edges[m] for m in member)
        print(json.dumps({
            "config": "BASELINE.json configs[4]: GraphSage_meanAggr 6x512 inference over a synthetic stiffened set, mesh "
                      "node counts x1 / x2 / x4 / x8 mixed, graph-sharded by greedy LPT (dist.partition_graphs), "
                      "device-side collate, batches cut by node budget",
            "n_gpus": world, "graphs": args.graphs, "distinct_meshes": args.pool, "nodes": tot_nodes, "edges": tot_edges,
            "max_hub_degree": max(nodes) - 1, "node_budget": args.node_budget, "precision": args.precision,
            "batches_rank0": len(batches), "seconds": sec, "graphs_per_s": args.graphs / sec,
            "nodes_per_s": tot_nodes / sec, "per_rank_ms": per_rank_ms,
            "lpt_imbalance_max_over_mean": max(loads) / (sum(loads) / world),
            "graphs_done": int(done.item()), "all_finite": bool(finite.item()),
            "same_mesh_same_prediction_rank0": bool(consistent)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
