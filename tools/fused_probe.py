"""Fused SAGE layer at growing sizes: correctness against aggregate-then-GEMM and per-launch time (tools probe)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buckgnn_b200 import engine
from buckgnn_b200.engine import Activation, build_graph_index
from buckgnn_b200.model import BuckGNN
from buckgnn_b200.synth import make_batch

DEV = "cuda:0"
torch.manual_seed(0)
cfg = dict(num_node_features=16, num_edge_features=5, hidden_channels=512, num_layers=2, pooling_layer="mean",
           model_name="GraphSage_meanAggr")
ours = BuckGNN(**cfg, precision="fp16").to(DEV).eval()
layer = ours._packed()["layers"][1]
sizes = [int(a) for a in sys.argv[1:]] or [2, 12, 40, 128, 256]
for g in sizes:
    b = make_batch(g, nx=64, ny=64)
    n = b.num_nodes
    idx = build_graph_index(b.edge_index.to(DEV), b.batch.to(DEV), n)
    x, agg, out_a, out_b = (Activation(n, 512, "fp16", DEV) for _ in range(4))
    x.data.copy_(torch.randn(n, 512, device=DEV).abs())
    print(f"graphs {g} nodes {n} tiles {(n + 255) // 256} n_big {idx.n_big}", flush=True)

    def unfused():
        engine.aggregate(x, agg, idx, "mean")
        segs = engine._segments(agg, layer.lin_l) + engine._segments(x, layer.lin_r)
        engine.gemm512(segs, n, "fp16", out_a, bias=layer.bias.data_ptr(), bn_scale=engine._p(layer.bn_scale),
                       bn_shift=engine._p(layer.bn_shift), residual=x.data.data_ptr(), ldr=512, normalize=True, relu=True)

    def fused():
        engine.sage_layer_fused(x, out_b, idx, layer, aggr="mean", relu=True, residual=True)

    for name, fn in (("unfused", unfused), ("fused", fused)):
        fn(); torch.cuda.synchronize()
        print(f"  {name} ran", flush=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record(); torch.cuda.synchronize()
        print(f"  {name}: {e0.elapsed_time(e1) / 5:.3f} ms per layer", flush=True)
    bad = (out_a.data != out_b.data).any(dim=1)
    print(f"  rows that differ: {int(bad.sum())} (hub rows differ by summation order: n_big = {idx.n_big}); "
          f"max abs diff {(out_a.data.float() - out_b.data.float()).abs().max().item():.3e}", flush=True)
