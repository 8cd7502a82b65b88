"""Per-call durations of the aggregation and SAGE-update launches of a forward, in layer order (cfg 2), with the
pool-fused last layer and without: is the last GEMM slower because of its epilogue or because of its position?"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buckgnn_b200 import engine
from buckgnn_b200.synth import config_batch
from tools.bench_configs import seeded_model

dev = "cuda:0"
b = config_batch(1).to(dev)
cfg = dict(num_node_features=16, num_edge_features=5, hidden_channels=512, num_layers=6, pooling_layer="mean",
           model_name="GraphSage_meanAggr")
for fuse in (True, False, True, False):
    model = seeded_model(cfg, "fp16", device=dev, fuse_pool=fuse)
    with torch.no_grad():
        for _ in range(3):
            model(b.x, b.edge_index, b.edge_attr, b.batch)
        torch.cuda.synchronize()
        engine.TIMERS.enable()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        steps = 20
        for _ in range(steps):
            model(b.x, b.edge_index, b.edge_attr, b.batch)
        e1.record()
        torch.cuda.synchronize()
    spans = list(engine.TIMERS.spans)
    engine.TIMERS.disable()
    per = {}
    order = []
    k = 0
    for name, s0, s1 in spans:
        per.setdefault(name, []).append(s0.elapsed_time(s1))
    line = f"fuse_pool={fuse}: step {e0.elapsed_time(e1) / steps:.3f} ms |"
    for name in ("aggregate", "sage_update", "sage_update_pool", "pool_head"):
        if name in per:
            v = per[name]
            n = len(v) // steps
            avg = [sum(v[i::n]) / steps for i in range(n)]
            line += f" {name}: " + " ".join(f"{a:.3f}" for a in avg) + " |"
    print(line, flush=True)
