"""N-GPU check of the sharded training step (torchrun): every rank runs its own graphs, gradients are
all-reduced (mean) with buckgnn_b200.dist.allreduce_gradients; the result must equal, on every rank, the mean
of the per-rank gradients that rank 0 recomputes by running every shard itself -- and parameters must stay
identical across ranks after an optimizer step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F
from buckgnn_b200 import train
from buckgnn_b200.dist import allreduce_gradients
from buckgnn_b200.model import BuckGNN
from buckgnn_b200.synth import config_batch


def main():
    world, rank, local = (int(os.environ[k]) for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = BuckGNN(16, 5, 512, 3, "mean", model_name="GraphSage_meanAggr", dropout_rate=0.1).to(dev).train()
    params = train.trainable_parameters(model)

    def shard_grads(r):
        b = config_batch(3, rank=r, num_graphs=4).to(dev)
        model.zero_grad(set_to_none=True)
        pred = train.forward_train(model, b.x, b.edge_index, b.batch, seed=1000 + r).squeeze()
        F.mse_loss(pred, b.y.to(dev)).backward()
        return [p.grad.clone() for p in params]

    bn_state = {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "tracked" in k}
    want = None
    if rank == 0:                                        # reference: all shards on one GPU, averaged
        per = [shard_grads(r) for r in range(world)]
        want = [sum(g[i] for g in per) / world for i in range(len(params))]
        model.load_state_dict(bn_state, strict=False)    # undo the extra BN buffer updates
    shard_grads(rank)
    n = allreduce_gradients(params)
    ok = torch.ones(1, device=dev)
    if rank == 0:
        err = max(((p.grad - w).norm() / w.norm().clamp(min=1e-20)).item() for p, w in zip(params, want))
        print(f"all-reduced {n} gradient elements over {world} ranks; max rel. deviation from the single-GPU mean: {err:.2e}")
        ok[0] = float(err < 1e-5)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool((lo == hi).all())
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"parameters identical on all ranks after the optimizer step: {same}")
        print("PASS" if same and ok.item() == 1 else "FAIL")
    dist.destroy_process_group()
    sys.exit(0 if same and ok.item() == 1 else 1)


if __name__ == "__main__":
    main()
