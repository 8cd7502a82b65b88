"""Throughput of the other BASELINE.json configs on one B200 (not the headline bench line):
  cfg3  stiffened-plate EA-GNN ("CustomGNN"), batch 128, fp16 operands / fp32 accumulate
  cfg5  GraphSAGE 6x512 on stiffened plates with mesh sizes scaled 1x..8x (hub degree = graph size)
  sag   the SAGPooling variants: GraphSAGE_SAG on the cfg2 batch (256 plates), EAGNN_SAG on 64 stiffened plates
Prints one JSON line per case."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buckgnn_b200 import capi, engine
from buckgnn_b200.model import BuckGNN
from buckgnn_b200.synth import make_batch

DEV = "cuda:0"


def timed(model, b, steps=5, warmup=2):
    with torch.no_grad():
        for _ in range(warmup):
            model(b.x, b.edge_index, b.edge_attr, b.batch)
        torch.cuda.synchronize()
        engine.TIMERS.enable()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            pred, _ = model(b.x, b.edge_index, b.edge_attr, b.batch)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    k = {n: v[0] / steps for n, v in engine.TIMERS.summary().items()}
    engine.TIMERS.disable()
    return ms, k, pred


def seeded_model(cfg, precision, device=None, **kw):
    """Seeded weights with BatchNorm running statistics of a trained network's magnitude (parity of these
    configurations against the oracle is tests/test_gpu_forward.py's job, not this tool's)."""
    torch.manual_seed(0)
    m = BuckGNN(**cfg, precision=precision, **kw)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                c = mod.num_features
                mod.running_mean.copy_(torch.randn(c, generator=g) * 0.3 / c ** 0.5)
                mod.running_var.copy_((torch.rand(c, generator=g) + 0.5) / c)
                mod.weight.copy_(torch.rand(c, generator=g) + 0.5)
                mod.bias.copy_(torch.randn(c, generator=g) * 0.1)
    return m.to(device or DEV).eval()


def main():
    capi.device_check()
    which = sys.argv[1:] or ["cfg3", "cfg5"]
    if "cfg3" in which:
        cfg = dict(num_node_features=16, num_edge_features=5, hidden_channels=512, num_layers=6,
                   pooling_layer="mean", model_name="EA_GNN")
        model = seeded_model(cfg, "fp16")
        b = make_batch(128, stiffened=True).to(DEV)
        ms, k, _ = timed(model, b)
        n, e = b.num_nodes, b.num_edges
        flops = 6 * (3 * e + 8 * n) * 2.0 * 512 * 512 - 2.0 * e * 512 * 512     # last layer skips edge_mlp L2
        print(json.dumps({"config": "cfg3: EA_GNN 6x512, batch 128 stiffened plates (CBAR sides+diagonals, 13.33% virtual "
                                    "edges, super node)", "graphs": 128, "nodes": n, "edges": e, "precision": "fp16",
                          "ms_per_forward": ms, "graphs_per_s": 128 / (ms * 1e-3), "gemm_tflops_as_executed": flops / ms / 1e9,
                          "kernel_ms": k}), flush=True)
        del model, b
        torch.cuda.empty_cache()
    if "cfg5" in which:
        cfg = dict(num_node_features=16, num_edge_features=5, hidden_channels=512, num_layers=6,
                   pooling_layer="mean", model_name="GraphSage_meanAggr")
        model = seeded_model(cfg, "fp16")
        for scale, graphs in ((1.0, 128), (2 ** 0.5, 64), (2.0, 32), (2 * 2 ** 0.5, 16)):
            b = make_batch(graphs, stiffened=True, scale=scale).to(DEV)
            ms, k, _ = timed(model, b)
            print(json.dumps({"config": f"cfg5: GraphSage_meanAggr 6x512 on stiffened plates, linear mesh scale {scale:.3f} "
                                        f"(node count x{scale * scale:.0f})", "graphs": graphs, "nodes": b.num_nodes,
                              "edges": b.num_edges, "max_hub_degree": int((b.ptr[1:] - b.ptr[:-1]).max()) - 1,
                              "ms_per_forward": ms, "graphs_per_s": graphs / (ms * 1e-3),
                              "nodes_per_s": b.num_nodes / (ms * 1e-3), "kernel_ms": k}), flush=True)
            del b
            torch.cuda.empty_cache()
    if "sag" in which or "sag1" in which:
        cases = (("GraphSAGE_SAG", "fp32", 256, {}), ("GraphSAGE_SAG", "tf32", 256, {}),
                                       ("GraphSAGE_SAG", "fp16", 256, {}), ("EAGNN_SAG", "fp32", 64, dict(stiffened=True)),
                                       ("EAGNN_SAG", "tf32", 64, dict(stiffened=True)))
        for name, prec, graphs, kw in (cases[:1] if "sag1" in which else cases):
            cfg = dict(num_node_features=16, num_edge_features=5, hidden_channels=512, num_layers=6,
                       pooling_layer="mean", model_name=name)
            model = seeded_model(cfg, prec)
            b = make_batch(graphs, **kw).to(DEV)
            ms, k, _ = timed(model, b)
            lp = model.last_pool
            print(json.dumps({"config": f"sag: {name} 6x512 (3 layers, SAGPooling ratio 0.5, 3 layers), batch {graphs}",
                              "precision": prec, "graphs": graphs, "nodes": b.num_nodes, "edges": b.num_edges,
                              "pooled_nodes": lp.n_nodes, "pooled_edges": lp.n_edges, "ms_per_forward": ms,
                              "graphs_per_s": graphs / (ms * 1e-3), "kernel_ms": k}), flush=True)
            del model, b
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
