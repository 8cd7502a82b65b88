"""Prediction error of the narrow-hidden twins by precision mode on a ragged batch (probe behind
tests/test_gpu_guard.py::test_narrow_hidden_stays_inside_its_tensors)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buckgnn_b200.model import BuckGNN
from buckgnn_b200.synth import collate, make_plate_graph
from oracle.buckgnn_oracle import OracleBuckGNN, randomize_bn_stats

DEV = "cuda:0"
sizes = [(3, 2), (31, 17), (2, 2), (40, 33), (9, 7), (23, 29)]
b = collate([make_plate_graph(i, nx=nx, ny=ny) for i, (nx, ny) in enumerate(sizes)])
bd = b.to(DEV)
for hidden in (64, 128, 256, 512):
    for layers in (3, 6):
        for seed in (0, 1, 2):
            torch.manual_seed(seed)
            cfg = dict(num_node_features=16, num_edge_features=5, hidden_channels=hidden, num_layers=layers,
                       pooling_layer="mean", model_name="GraphSage_meanAggr")
            ref = OracleBuckGNN(**cfg).eval()
            randomize_bn_stats(ref, realistic=True)
            ref64 = OracleBuckGNN(**cfg).double().eval()
            ref64.load_state_dict(ref.state_dict())
            with torch.no_grad():
                want, _ = ref(b.x, b.edge_index, b.edge_attr, b.batch)
                want64, _ = ref64(b.x.double(), b.edge_index, b.edge_attr.double(), b.batch)
            row = [f"h={hidden} L={layers} seed={seed} | fp32-oracle vs fp64 {((want.double()-want64).abs()/want64.abs().clamp(min=1e-3)).max().item():.1e}"]
            for prec in ("fp16", "bf16", "tf32", "fp32"):
                ours = BuckGNN(**cfg, precision=prec)
                ours.load_state_dict(ref.state_dict())
                ours = ours.to(DEV).eval()
                with torch.no_grad():
                    got, _ = ours(bd.x, bd.edge_index, bd.edge_attr, bd.batch)
                rel = ((got.cpu().double() - want64).abs() / want64.abs().clamp(min=1e-3))
                row.append(f"{prec} {rel.max().item():.1e} (graph {int(rel.argmax())})")
            print(" | ".join(row), flush=True)
