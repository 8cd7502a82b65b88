"""Training-step throughput, BASELINE.json configs[3]: TRAIN_FINAL-style step of GraphSage_meanAggr 6x512,
batch 16 graphs per GPU, graph-sharded, gradient all-reduce over NCCL (one flat bucket).

    python tools/bench_train.py [--graphs 16] [--steps 20] [--precision tf32|bf16]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py

One step = zero_grad -> model(...) in train mode (batch-statistics BatchNorm, dropout 0.1) -> relative-error
loss -> loss.backward() (backward kernels) -> all-reduce of the gradients -> Adam step, as
TRAIN_FINAL.py:289-297.  Prints one JSON line (rank 0): graphs/s over all ranks (max-over-ranks step time,
CUDA events), the per-kernel-class milliseconds of one step, and the all-reduce time."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from buckgnn_b200 import capi, engine, train
from buckgnn_b200.dist import allreduce_gradients
from buckgnn_b200.model import BuckGNN
from buckgnn_b200.synth import config_batch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graphs", type=int, default=16)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--model", default="GraphSage_meanAggr", help="GraphSage_*Aggr | GraphSAGE_SAG | EA_GNN | EA_GNN_Shared | EAGNN_SAG")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    capi.device_check()
    torch.manual_seed(0)                                   # same initial weights on every rank
    model = BuckGNN(16, 5, 512, 6, "mean", model_name=args.model, dropout_rate=0.1,
                    train_precision=args.precision).to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    stiff = args.model in ("EA_GNN", "EA_GNN_Shared", "EAGNN_SAG")   # the EA-GNN configs use the stiffened plates (cfg 3)
    b = config_batch(2 if stiff else 3, rank=rank, num_graphs=args.graphs).to(dev)
    y = b.y.to(dev).abs() + 0.5
    params = train.trainable_parameters(model)
    ar0, ar1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from buckgnn_b200.loss import EigenvalueRelativeLoss
    crit = EigenvalueRelativeLoss(scale=1.0, center=0.0)

    def step():
        opt.zero_grad(set_to_none=True)
        pred, _ = model(b.x, b.edge_index, b.edge_attr, b.batch)
        loss = crit(pred, y)                                # RelativeErrorLoss + MAPE, fused (buckgnn_b200/loss.py)
        loss.backward()
        ar0.record()
        allreduce_gradients(params)
        ar1.record()
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    engine.TIMERS.enable()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ar_ms = 0.0
    losses = []
    for _ in range(args.steps):
        losses.append(step())
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ar_ms = ar0.elapsed_time(ar1)
    ms = e0.elapsed_time(e1) / args.steps
    k = {n: round(v[0] / args.steps, 4) for n, v in engine.TIMERS.summary().items()}
    engine.TIMERS.disable()
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    if rank == 0:
        n_grad = sum(p.numel() for p in params)
        print(json.dumps({
            "config": f"cfg4: training step, {args.model} 6x512, dropout 0.1, Adam, batch "
                      f"{args.graphs} graphs per GPU, graph-sharded x{world}, gradient all-reduce (NCCL, one flat bucket)",
            "n_gpus": world, "graphs_per_gpu": args.graphs, "nodes_per_gpu": b.num_nodes, "edges_per_gpu": b.num_edges,
            "train_precision": args.precision, "ms_per_step": ms, "graphs_per_s": world * args.graphs / (ms * 1e-3),
            "allreduce_ms_last_step": ar_ms, "allreduce_elements": n_grad,
            "kernel_ms_per_step": k, "loss_first": float(losses[0].detach()), "loss_last": float(losses[-1].detach())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
