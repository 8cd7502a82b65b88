"""bg_sage_aggregate alone on a STIFFENED batch (configs[2] / configs[4] meshes: degree ~11 with the CBAR diagonals
and virtual edges): ms per launch and algorithmic GB/s; the ncu target for the degree-11 case."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buckgnn_b200 import engine
from buckgnn_b200.engine import Activation
from buckgnn_b200.synth import make_batch

dev = "cuda:0"
b = make_batch(128, stiffened=True)
n = b.num_nodes
idx = engine.build_graph_index(b.edge_index.to(dev), b.batch.to(dev), n)
x, o = Activation(n, 512, "fp16", dev), Activation(n, 512, "fp16", dev)
x.data.copy_(torch.randn(n, 512, device=dev))
alg = 2 * n * 512 * 2 + 4 * idx.n_edges + 4 * (n + 1)
for _ in range(3):
    engine.aggregate(x, o, idx, "mean")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    engine.aggregate(x, o, idx, "mean")
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"stiffened x128: N={n} E={idx.n_edges} n_big={idx.n_big} range_hubs={idx.hub_lo is not None}  {ms:.4f} ms  "
      f"{alg / ms / 1e6:.0f} GB/s algorithmic, {idx.n_edges * 1024 / ms / 1e6:.0f} GB/s gathered")
