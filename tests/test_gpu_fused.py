"""The fused SAGE layer (bg_sage_fused512): the aggregate operand gathered inside the update GEMM must be bit-identical
to the two-kernel layer, whose aggregate matrix went through global memory."""
import pytest
import torch

from buckgnn_b200 import engine
from buckgnn_b200.engine import Activation, build_graph_index
from buckgnn_b200.model import BuckGNN
from buckgnn_b200.synth import collate, make_batch, make_plate_graph
from oracle.buckgnn_oracle import OracleBuckGNN, randomize_bn_stats

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pair(name, precision, layers=4, seed=0):
    torch.manual_seed(seed)
    cfg = dict(num_node_features=16, num_edge_features=5, hidden_channels=512, num_layers=layers,
               pooling_layer="mean", model_name=name)
    ref = OracleBuckGNN(**cfg).eval()
    randomize_bn_stats(ref, realistic=True)
    ours = BuckGNN(**cfg, precision=precision)
    ours.load_state_dict(ref.state_dict())
    return ref, ours.to(DEV).eval()


def _rel(got, want):
    return ((got - want).abs() / want.abs().clamp(min=1e-3)).max().item()


def _batches():
    ragged = collate([make_plate_graph(i, nx=nx, ny=ny) for i, (nx, ny) in
                      enumerate([(3, 2), (31, 17), (2, 2), (40, 33), (9, 7), (23, 29)])])
    return {"mesh": make_batch(4, nx=24, ny=20), "ragged": ragged,
            "stiffened": make_batch(3, nx=16, ny=12, stiffened=True)}           # degree ~11: more than 8 neighbours per row


@pytest.mark.parametrize("case", ["mesh", "ragged", "stiffened"])
@pytest.mark.parametrize("precision", ["fp16", "bf16"])
@pytest.mark.parametrize("aggr,residual,relu", [("mean", True, True), ("mean", False, True), ("sum", True, False)])
def test_fused_layer_is_bit_identical_to_aggregate_then_gemm(case, precision, aggr, residual, relu):
    b = _batches()[case]
    n = b.num_nodes
    _, ours = _pair("GraphSage_meanAggr", precision, layers=2)
    layer = ours._packed()["layers"][1]
    idx = build_graph_index(b.edge_index.to(DEV), b.batch.to(DEV), n)
    assert idx.n_big >= 1
    torch.manual_seed(3)
    x, agg, out_a, out_b = (Activation(n, 512, precision, DEV) for _ in range(4))
    x.data.copy_(torch.randn(n, 512).abs())
    out_a.data.fill_(7.0); out_b.data.fill_(7.0)
    # two kernels; hub rows through the generic hub kernel, as the fused path's side buffer is
    engine.aggregate(x, agg, idx, aggr, fold_hubs=False)
    segs = engine._segments(agg, layer.lin_l) + engine._segments(x, layer.lin_r)
    engine.gemm512(segs, n, precision, out_a, bias=layer.bias.data_ptr(), bn_scale=engine._p(layer.bn_scale),
                   bn_shift=engine._p(layer.bn_shift), residual=x.data.data_ptr() if residual else None, ldr=512,
                   normalize=True, relu=relu)
    engine.sage_layer_fused(x, out_b, idx, layer, aggr=aggr, relu=relu, residual=residual)
    torch.cuda.synchronize()
    assert torch.isfinite(out_b.data.float()).all()
    assert torch.equal(out_a.data, out_b.data)
    # and again: deterministic
    out_c = Activation(n, 512, precision, DEV)
    engine.sage_layer_fused(x, out_c, idx, layer, aggr=aggr, relu=relu, residual=residual)
    assert torch.equal(out_c.data, out_b.data)


@pytest.mark.parametrize("case", ["mesh", "ragged", "stiffened"])
@pytest.mark.parametrize("name,precision", [("GraphSage_meanAggr", "fp16"), ("GraphSage_meanAggr", "bf16"),
                                            ("GraphSage_sumAggr", "fp16")])
def test_fused_forward_matches_unfused_and_oracle(case, name, precision):
    b = _batches()[case]
    ref, ours = _pair(name, precision)
    bd = b.to(DEV)
    with torch.no_grad():
        want, _ = ref(b.x, b.edge_index, b.edge_attr, b.batch)
        ours.fuse_aggregate = False
        plain, _ = ours(bd.x, bd.edge_index, bd.edge_attr, bd.batch)
        ours.fuse_aggregate = True
        fused, _ = ours(bd.x, bd.edge_index, bd.edge_attr, bd.batch)
    plain, fused = plain.float().cpu(), fused.float().cpu()
    assert torch.isfinite(fused).all()
    assert _rel(fused, plain) < 2e-4          # (hub rows: range fold vs generic hub kernel = another summation order)
    if case != "ragged" and precision == "fp16":
        assert _rel(fused, want) < 1e-3


def test_fused_layer_rejects_what_it_cannot_do():
    from buckgnn_b200 import capi
    b = make_batch(2, nx=6, ny=5)
    n = b.num_nodes
    _, ours = _pair("GraphSage_meanAggr", "tf32", layers=2)
    layer = ours._packed()["layers"][1]
    idx = build_graph_index(b.edge_index.to(DEV), b.batch.to(DEV), n)
    x, out = Activation(n, 512, "tf32", DEV), Activation(n, 512, "tf32", DEV)
    assert not engine.can_fuse_aggregate("tf32", "mean", True) and not engine.can_fuse_aggregate("fp16", "max", True)
    with pytest.raises(capi.BuckGNNError):
        engine.sage_layer_fused(x, out, idx, layer, aggr="mean", relu=True, residual=False)


def test_fused_layer_with_several_tiles_per_cta_and_repeated_launches():
    """193 row tiles on 74 CTA pairs: every pair walks 2-3 tiles, so the gather warps cross tile boundaries of the
    4-slot operand ring (the round-2 bug: a parity wait that skipped the slots filled by the producer alone took the
    wrong phase for its own and overwrote tiles the tensor core had not read -- only when the gathers were FAST, i.e.
    from the second launch on, with the rows in L2)."""
    b = make_batch(12, nx=64, ny=64)
    n = b.num_nodes
    _, ours = _pair("GraphSage_meanAggr", "fp16", layers=2)
    layer = ours._packed()["layers"][1]
    idx = build_graph_index(b.edge_index.to(DEV), b.batch.to(DEV), n)
    x, agg, out_a = (Activation(n, 512, "fp16", DEV) for _ in range(3))
    x.data.copy_(torch.randn(n, 512, device=DEV).abs())
    engine.aggregate(x, agg, idx, "mean", fold_hubs=False)
    segs = engine._segments(agg, layer.lin_l) + engine._segments(x, layer.lin_r)
    engine.gemm512(segs, n, "fp16", out_a, bias=layer.bias.data_ptr(), bn_scale=engine._p(layer.bn_scale),
                   bn_shift=engine._p(layer.bn_shift), residual=x.data.data_ptr(), ldr=512, normalize=True, relu=True)
    for _ in range(6):
        out_b = Activation(n, 512, "fp16", DEV)
        engine.sage_layer_fused(x, out_b, idx, layer, aggr="mean", relu=True, residual=True)
        assert torch.equal(out_a.data, out_b.data)
