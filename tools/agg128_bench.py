"""GPU micro-benchmark of the 128-column aggregation of the folded layer 0 (bg_sage_aggregate, width 128) on the cfg-2
batch: ms per launch (row kernel + hub kernel) and algorithmic GB/s."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buckgnn_b200 import engine
from buckgnn_b200.engine import Activation
from buckgnn_b200.synth import config_batch

DEV = "cuda:0"


def main(iters=20, cfg=1):
    b = config_batch(cfg)
    n = b.num_nodes
    idx = engine.build_graph_index(b.edge_index.to(DEV), b.batch.to(DEV), n)
    for precision in ("fp16",):
        x, o = Activation(n, 128, precision, DEV), Activation(n, 128, precision, DEV)
        x.data.copy_(torch.relu(torch.randn(n, 128, device=DEV)))
        es = x.data.element_size()
        alg = 2 * n * 128 * es + 4 * idx.n_edges + 4 * (n + 1)
        for _ in range(3):
            engine.aggregate(x, o, idx, "mean")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            engine.aggregate(x, o, idx, "mean")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"cfg{cfg} {precision} width 128: {ms:.4f} ms  {alg / ms / 1e6:.0f} GB/s algorithmic", flush=True)


if __name__ == "__main__":
    main(iters=int(sys.argv[1]) if len(sys.argv) > 1 else 20)
