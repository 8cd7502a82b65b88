"""Round-2 parity cases: narrow hidden widths on the padded twin, configs[4]'s
model on stiffened + virtual-edge meshes (generic hub rows included), exact super-node degrees in the folded first
layer, BatchNorm statistics after a train-mode forward, finite-value guard of the fp16 path."""
import pytest
import torch
import torch.nn.functional as F

from buckgnn_b200.model import BuckGNN
from buckgnn_b200.synth import PlateBatch, make_batch
from oracle.buckgnn_oracle import OracleBuckGNN, randomize_bn_stats

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cfg(**over):
    cfg = dict(num_node_features=16, num_edge_features=5, hidden_channels=512, num_layers=6, pooling_layer="mean",
               model_name="GraphSage_meanAggr")
    cfg.update(over)
    return cfg


def _pair(cfg, seed=0, **kw):
    torch.manual_seed(seed)
    ref = OracleBuckGNN(**cfg).eval()
    randomize_bn_stats(ref, realistic=True)
    ours = BuckGNN(**cfg, **kw)
    ours.load_state_dict(ref.state_dict())
    return ref, ours.to(DEV).eval()


def _fwd(ref, ours, b):
    with torch.no_grad():
        want, _ = ref(b.x, b.edge_index, b.edge_attr, b.batch)
        bd = b.to(DEV)
        got, _ = ours(bd.x, bd.edge_index, bd.edge_attr, bd.batch)
    return got.cpu(), want


def _rel(got, want):
    return ((got - want).abs() / want.abs().clamp(min=1e-3)).max().item()


# ----------------------------------------------------------------------------- hidden_channels != 512 (TRAIN_FINAL.py:55,71)
@pytest.mark.parametrize("hidden", [64, 128, 256])
@pytest.mark.parametrize("name,precision", [("GraphSage_meanAggr", "fp16"), ("GraphSage_addAggr_Shared", "tf32"),
                                            ("EA_GNN", "tf32"), ("GraphSAGE_SAG", "fp32")])
def test_narrow_hidden_forward_matches_oracle(name, precision, hidden):
    ref, ours = _pair(_cfg(model_name=name, hidden_channels=hidden, num_layers=4), precision=precision)
    b = make_batch(3, nx=12, ny=10, stiffened=name == "EA_GNN")
    got, want = _fwd(ref, ours, b)
    assert got.shape == want.shape == (3,)
    assert _rel(got, want) < 1e-3


@pytest.mark.parametrize("pooling", ["supernode_with_pooling", "mlp", "supernode_only"])
def test_narrow_hidden_poolings_and_node_heads(pooling):
    """The padded twin's layouts (cat decoder blocks, MLPPooling, node-level decoder) in the fp32-GEMM mode, at its
    1e-4 bar: a single super-node row or per-node outputs do not average operand rounding over a graph."""
    ref, ours = _pair(_cfg(hidden_channels=128, num_layers=3, pooling_layer=pooling), precision="fp32")
    got, want = _fwd(ref, ours, make_batch(3, nx=9, ny=8))
    assert _rel(got, want) < 1e-4
    ref, ours = _pair(_cfg(hidden_channels=128, num_layers=3, pooling_layer=pooling, prediction_type="static_stress"),
                      precision="fp32")
    got, want = _fwd(ref, ours, make_batch(2, nx=9, ny=8))
    assert got.shape == want.shape
    assert ((got.double() - want.double()).norm() / want.double().norm()).item() < 1e-4


def test_narrow_hidden_training_step_like_train_final():
    """TRAIN_FINAL.py:289-297 at its own width (128): loss, gradients on the NARROW parameters, BatchNorm buffers."""
    cfg = _cfg(hidden_channels=128, num_layers=3, dropout_rate=0.0)
    torch.manual_seed(0)
    ref = OracleBuckGNN(**cfg)
    randomize_bn_stats(ref, realistic=True)
    ours = BuckGNN(**cfg, train_precision="tf32")
    ours.load_state_dict(ref.state_dict())
    ref, ours = ref.train(), ours.to(DEV).train()
    b = make_batch(4, nx=12, ny=10)
    y = torch.tensor([0.5, -0.25, 1.0, 0.1])
    pw, _ = ref(b.x, b.edge_index, b.edge_attr, b.batch)
    lw = F.mse_loss(pw, y)
    lw.backward()
    bd = b.to(DEV)
    pg, bb = ours(bd.x, bd.edge_index, bd.edge_attr, bd.batch)
    assert bb is bd.batch and pg.shape == (4,)
    lg = F.mse_loss(pg, y.to(DEV))
    lg.backward()
    assert abs(lg.item() - lw.item()) < 2e-3 * abs(lw.item())
    rp, op = dict(ref.named_parameters()), dict(ours.named_parameters())
    checked = 0
    for k, p in rp.items():
        if p.grad is None:
            assert op[k].grad is None, k
            continue
        assert op[k].grad is not None and op[k].grad.shape == p.shape, k
        if p.grad.norm() > 1e-10:
            # independent ReLU masks on the two sides: a few % of flip noise (tests/test_gpu_train.py explains)
            err = ((op[k].grad.cpu().double() - p.grad.double()).norm() / p.grad.double().norm()).item()
            assert err < 1.5e-1, (k, err)
            checked += 1
    assert checked >= 10
    for (k, rb), (_, ob) in zip(ref.named_buffers(), ours.named_buffers()):
        if rb.dtype.is_floating_point:
            assert ((ob.cpu().double() - rb.double()).norm() / rb.double().norm()).item() < 2e-3, k
        else:
            assert int(ob) == int(rb), k
    opt = torch.optim.Adam(ours.parameters(), lr=1e-3)
    losses = []
    for _ in range(4):
        opt.zero_grad()
        pred, _ = ours(bd.x, bd.edge_index, bd.edge_attr, bd.batch)
        loss = F.mse_loss(pred, y.to(DEV))
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]


def test_default_model_name_trains():
    """model_name='GraphSAGE_MLP' (the constructor default, Models/BuckGNN.py:12) matches no branch: encoder -> pool ->
    decoder.  It has a gradient like any other."""
    cfg = _cfg(model_name="GraphSAGE_MLP", hidden_channels=128, dropout_rate=0.0)
    torch.manual_seed(0)
    ref = OracleBuckGNN(**cfg).train()
    ours = BuckGNN(**cfg)
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(DEV).train()
    b = make_batch(3, nx=8, ny=7)
    y = torch.tensor([0.3, -0.2, 0.9])
    pw, _ = ref(b.x, b.edge_index, b.edge_attr, b.batch)
    F.mse_loss(pw, y).backward()
    bd = b.to(DEV)
    pg, _ = ours(bd.x, bd.edge_index, bd.edge_attr, bd.batch)
    F.mse_loss(pg, y.to(DEV)).backward()
    assert _rel(pg.detach().cpu(), pw.detach()) < 1e-3
    g_ref, g_our = ref.decoder[0].weight.grad, ours.decoder[0].weight.grad.cpu()
    assert ((g_our - g_ref).norm() / g_ref.norm()).item() < 5e-2


# ----------------------------------------------------------------------------- configs[4]'s model on its meshes
def _shuffle_nodes(b: PlateBatch, seed: int) -> PlateBatch:
    """Relabels the nodes inside every graph at random: the super node is no longer the last row and its neighbour
    list is no contiguous run -> the hub rows take the generic (non-range) hub path."""
    g = torch.Generator().manual_seed(seed)
    perm = torch.empty(b.num_nodes, dtype=torch.int64)
    for i in range(b.num_graphs):
        lo, hi = int(b.ptr[i]), int(b.ptr[i + 1])
        perm[lo:hi] = lo + torch.randperm(hi - lo, generator=g)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(b.num_nodes)
    return PlateBatch(b.x[perm].contiguous(), inv[b.edge_index].contiguous(), b.edge_attr, b.batch, b.y, b.ptr, b.num_graphs)


@pytest.mark.parametrize("precision", ["fp16", "tf32"])
@pytest.mark.parametrize("layout", ["stiffened_virtual", "no_super_virtual", "shuffled_nodes"])
def test_meanaggr_on_stiffened_virtual_edge_meshes(layout, precision):
    ref, ours = _pair(_cfg(), precision=precision)
    if layout == "stiffened_virtual":
        b = make_batch(3, nx=40, ny=36, stiffened=True)                       # hub degree 1440: a "big row"
    elif layout == "no_super_virtual":
        b = make_batch(3, nx=20, ny=18, stiffened=True, super_node=False, virtual_edges=True)
    else:
        b = _shuffle_nodes(make_batch(3, nx=40, ny=36, stiffened=True), seed=5)
    got, want = _fwd(ref, ours, b)
    assert _rel(got, want) < 1e-3


# ----------------------------------------------------------------------------- ADVICE r01
@pytest.mark.parametrize("name", ["GraphSage_addAggr", "GraphSage_sumAggr"])
@pytest.mark.parametrize("precision,rtol", [("fp32", 1e-4), ("tf32", 1e-3), ("fp16", 1e-3)])
def test_folded_layer0_counts_a_super_node_degree_exactly(name, precision, rtol):
    """sum / add aggregation: the folded first layer multiplies W_l b3 by the in-degree.  A super node of a 50 x 45 plate has
    degree 2250 -- not representable in the 11 (8) significant bits of a tf32 / fp16 (bf16) operand; the degree
    travels as base-256 digits instead."""
    ref, ours = _pair(_cfg(model_name=name, num_layers=2), precision=precision)
    b = make_batch(2, nx=50, ny=45)
    got, want = _fwd(ref, ours, b)
    _, unfolded = _pair(_cfg(model_name=name, num_layers=2), precision=precision, fold_encoder=False)
    got_u, _ = _fwd(ref, unfolded, b)
    assert _rel(got, want) < rtol
    assert _rel(got, got_u) < rtol


def test_train_forward_without_optimizer_step_then_eval():
    """The train-mode kernels update running_mean / running_var through raw pointers; the eval-mode operand cache must
    not keep the BatchNorm fold of the OLD statistics (BN recalibration, a GradScaler-skipped step, train() -> eval())."""
    cfg = _cfg(num_layers=3, dropout_rate=0.0)
    ref, ours = _pair(cfg, precision="tf32")
    b = make_batch(3, nx=12, ny=10)
    got0, want0 = _fwd(ref, ours, b)                         # builds the eval cache
    assert _rel(got0, want0) < 1e-3
    ref.train(); ours.train()
    bd = b.to(DEV)
    with torch.no_grad():
        ref(b.x, b.edge_index, b.edge_attr, b.batch)
        ours(bd.x, bd.edge_index, bd.edge_attr, bd.batch)
    ref.eval(); ours.eval()
    got1, want1 = _fwd(ref, ours, b)
    assert (want1 - want0).abs().max().item() > 1e-3         # the statistics really moved
    assert _rel(got1, want1) < 1e-3


def test_fp16_mode_reports_non_finite_activations():
    """fp16 storage relies on L2-normalize + BatchNorm keeping activations small; the encoder's hidden layer is
    stored before any normalisation.  Weights that push it past 65504 must raise, not return inf silently."""
    ref, ours = _pair(_cfg(num_layers=2), precision="fp16")
    with torch.no_grad():
        ours.node_encoder[2].weight.mul_(1e5)
        ours.node_encoder[2].bias.fill_(1e5)
    bd = make_batch(2, nx=8, ny=7).to(DEV)
    with torch.no_grad():
        pred, _ = ours(bd.x, bd.edge_index, bd.edge_attr, bd.batch)
    assert torch.isnan(pred).all()                           # loud in the result, without a host sync in the forward
    with pytest.raises(FloatingPointError):
        ours.check_finite()
    _, ok = _pair(_cfg(num_layers=2), precision="fp16")
    with torch.no_grad():
        pred, _ = ok(bd.x, bd.edge_index, bd.edge_attr, bd.batch)
    assert torch.isfinite(pred).all()
    ok.check_finite()
