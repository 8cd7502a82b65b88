"""configs[3] training step under torchrun with different gradient-bucket sizes (dist.GradSync.min_bucket_elems):
how much of the multi-GPU step is collective time and how much is the ranks waiting for each other at every collective.
    python -m torch.distributed.run --nproc-per-node N ... tools/train_bucket_probe.py"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


args = argparse.Namespace(steps=20, train_precision="tf32")
for elems in (1 << 20, 1 << 21, 1 << 23):
    os.environ["BUCKGNN_GRAD_BUCKET_ELEMS"] = str(elems)
    for rep in range(2):
        out = bench.train_section(args, dev, world, rank, barrier)
        if rank == 0:
            print(json.dumps({"bucket_elems": elems, "rep": rep, **{k: out[k] for k in out if k in (
                "ms_per_step", "graphs_per_s", "allreduce_ms", "allreduce_buckets", "ms_per_step_without_allreduce", "allreduce_exposed_ms")}}), flush=True)
if world > 1:
    dist.destroy_process_group()
