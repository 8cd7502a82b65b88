"""Generates tests/golden/*.json from the CPU oracle (oracle/buckgnn_oracle.py).

    python tests/golden/make_golden.py

PyG / torch_scatter cannot be installed in this environment and the reference ships no golden
vectors, so these fixtures do not pin the oracle against the reference itself ("parity unpinned",
DESIGN.md section 3); they FREEZE the oracle -- hand-checked on the known-answer cases of
tests/test_oracle.py -- so that neither it nor the CUDA path can drift unnoticed.  Inputs are
regenerated from seeds (buckgnn_b200.synth, torch.manual_seed); outputs are stored.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

from buckgnn_b200.synth import make_batch
from oracle import buckgnn_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))

FORWARD_CASES = [  # name, model cfg overrides, batch kwargs
    ("sage_mean_6x512", dict(model_name="GraphSage_meanAggr"), dict(num_graphs=3, nx=9, ny=7)),
    ("sage_sum_4x512", dict(model_name="GraphSage_sumAggr", num_layers=4), dict(num_graphs=2, nx=8, ny=6)),
    ("sage_max_4x512", dict(model_name="GraphSage_maxAggr", num_layers=4), dict(num_graphs=2, nx=8, ny=6)),
    ("sage_add_shared_4x512", dict(model_name="GraphSage_addAggr_Shared", num_layers=4), dict(num_graphs=2, nx=8, ny=6)),
    ("eagnn_3x512_stiffened", dict(model_name="EA_GNN", num_layers=3), dict(num_graphs=2, nx=7, ny=6, stiffened=True)),
    ("default_mlp", dict(model_name="GraphSAGE_MLP"), dict(num_graphs=3, nx=6, ny=5)),
    ("sage_mean_supernode_only", dict(model_name="GraphSage_meanAggr", num_layers=3, pooling_layer="supernode_only"),
     dict(num_graphs=3, nx=6, ny=5)),
]


def model_cfg(**over):
    cfg = dict(num_node_features=16, num_edge_features=5, hidden_channels=512, num_layers=6,
               pooling_layer="mean", model_name="GraphSage_meanAggr")
    cfg.update(over)
    return cfg


def seeded_oracle(cfg):
    torch.manual_seed(0)
    m = O.OracleBuckGNN(**cfg)
    O.randomize_bn_stats(m, realistic=True)
    return m


def state_checksum(m):
    """Order-sensitive fp64 checksum of the seeded parameters and buffers (detects RNG drift)."""
    s = 0.0
    for i, (_, t) in enumerate(m.state_dict().items()):
        s += (i + 1) * float(t.double().sum())
    return s


def kat():
    """The 5-node / 2-graph case of tests/test_oracle.py: an isolated node and a hub."""
    x = torch.arange(10, dtype=torch.float32).view(5, 2) / 4 - 1
    ei = torch.tensor([[0, 1, 3, 3, 0, 1], [2, 2, 2, 4, 1, 0]])
    batch = torch.tensor([0, 0, 0, 1, 1])
    return x, ei, batch


def main():
    out = {"forward": {}, "operators": {}, "training": {}}
    for name, over, bkw in FORWARD_CASES:
        cfg = model_cfg(**over)
        m = seeded_oracle(cfg).eval()
        b = make_batch(**bkw)
        with torch.no_grad():
            pred, _ = m(b.x, b.edge_index, b.edge_attr, b.batch)
        out["forward"][name] = {"cfg": cfg, "batch": bkw, "nodes": b.num_nodes, "edges": b.num_edges,
                                "state_checksum": state_checksum(m), "pred": pred.double().reshape(-1).tolist()}
    x, ei, batch = kat()
    ops = out["operators"]
    for aggr in ("mean", "sum", "max"):
        ops[f"aggregate_{aggr}"] = O.aggregate(x, ei, aggr).tolist()
    ops["global_mean_pool"] = O.global_mean_pool(x, batch).tolist()
    ops["scatter_mean_row"] = O.scatter_mean(x[ei[1]], ei[0], 5).tolist()
    # training step: loss and gradient norms of one step (dropout off), plus BN buffers after it
    cfg = model_cfg(num_layers=3, dropout_rate=0.0)
    m = seeded_oracle(cfg).train()
    b = make_batch(num_graphs=3, nx=7, ny=6)
    y = torch.tensor([0.5, -0.25, 1.0])
    pred, _ = m(b.x, b.edge_index, b.edge_attr, b.batch)
    loss = torch.nn.functional.mse_loss(pred, y)
    loss.backward()
    out["training"]["sage_mean_3x512_step"] = {
        "cfg": cfg, "batch": dict(num_graphs=3, nx=7, ny=6), "y": y.tolist(), "loss": float(loss),
        "pred": pred.detach().double().tolist(),
        "grad_norms": {k: float(p.grad.double().norm()) for k, p in m.named_parameters() if p.grad is not None},
        "no_grad": sorted(k for k, p in m.named_parameters() if p.grad is None),
        "bn0_running_mean_sum": float(m.batch_norms[0].running_mean.double().sum()),
        "bn0_running_var_sum": float(m.batch_norms[0].running_var.double().sum())}
    path = os.path.join(HERE, "oracle_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
