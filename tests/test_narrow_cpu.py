"""hidden_channels != 512 (TRAIN_FINAL.py:55,71 use 128): the zero-padded 512-wide twin of buckgnn_b200/narrow.py is
EXACTLY the narrow model.  Checked on the CPU by loading the twin's parameters into a 512-wide oracle and comparing with
the narrow oracle (the padded columns must stay 0 through every layer), and by pushing a gradient back through the
embedding."""
import pytest
import torch

from buckgnn_b200.model import BuckGNN
from buckgnn_b200.narrow import WideTwin
from buckgnn_b200.synth import make_batch
from oracle import buckgnn_oracle as O

NAMES = ["GraphSage_meanAggr", "GraphSage_addAggr", "GraphSage_maxAggr", "GraphSage_addAggr_Shared", "EA_GNN",
         "EA_GNN_Shared", "GraphSAGE_SAG", "EAGNN_SAG", "GraphSAGE_MLP"]


def _pair(h, name, pooling="mean", ptype="buckling", layers=3):
    cfg = dict(num_node_features=16, num_edge_features=5, hidden_channels=h, num_layers=layers, pooling_layer=pooling,
               prediction_type=ptype, model_name=name, dropout_rate=0.0)
    torch.manual_seed(0)
    narrow_oracle = O.OracleBuckGNN(**cfg)
    O.randomize_bn_stats(narrow_oracle, realistic=True)
    ours = BuckGNN(**cfg)
    ours.load_state_dict(narrow_oracle.state_dict())
    twin = WideTwin(ours, ours._ctor_kwargs)
    twin.sync(differentiable=False)
    wide_oracle = O.OracleBuckGNN(**dict(cfg, hidden_channels=512))
    wide_oracle.load_state_dict(twin.twin.state_dict(), strict=True)
    return narrow_oracle.eval(), wide_oracle.eval(), ours, twin


@pytest.mark.parametrize("h", [64, 128, 256])
@pytest.mark.parametrize("name", NAMES)
def test_padded_twin_equals_narrow_model(name, h):
    stiff = name in ("EA_GNN", "EA_GNN_Shared", "EAGNN_SAG")
    no, wo, _, _ = _pair(h, name)
    b = make_batch(num_graphs=3, nx=5, ny=4, stiffened=stiff)
    with torch.no_grad():
        pn, _ = no(b.x, b.edge_index, b.edge_attr, b.batch)
        pw, _ = wo(b.x, b.edge_index, b.edge_attr, b.batch)
    torch.testing.assert_close(pw, pn, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("h", [128, 256])
@pytest.mark.parametrize("pooling", ["supernode_with_pooling", "mlp", "supernode_only"])
def test_padded_twin_poolings(pooling, h):
    no, wo, _, _ = _pair(h, "GraphSage_meanAggr", pooling=pooling)
    b = make_batch(num_graphs=2, nx=5, ny=4)
    with torch.no_grad():
        pn, _ = no(b.x, b.edge_index, b.edge_attr, b.batch)
        pw, _ = wo(b.x, b.edge_index, b.edge_attr, b.batch)
    torch.testing.assert_close(pw, pn, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("h", [128, 256])
def test_padded_twin_node_level_head(h):
    no, wo, _, _ = _pair(h, "GraphSage_meanAggr", ptype="static_stress")
    b = make_batch(num_graphs=2, nx=5, ny=4)
    with torch.no_grad():
        pn, _ = no(b.x, b.edge_index, b.edge_attr, b.batch)
        pw, _ = wo(b.x, b.edge_index, b.edge_attr, b.batch)
    torch.testing.assert_close(pw, pn, rtol=2e-5, atol=1e-6)


def test_gradients_slice_back_onto_the_narrow_parameters():
    _, _, ours, twin = _pair(128, "GraphSage_meanAggr")
    ours.train()
    live = twin.sync(differentiable=True)
    conv_t, conv_n = twin.twin.sage_blocks_mean[1], ours.sage_blocks_mean[1]
    g = torch.randn(512, 512)
    (live[id(conv_t.lin_l.weight)] * g).sum().backward()
    torch.testing.assert_close(conv_n.lin_l.weight.grad, g[:128, :128])
    dec_t, dec_n = twin.twin.decoder, ours.decoder
    g2 = torch.randn(128, 512)
    (live[id(dec_t[0].weight)] * g2).sum().backward()
    torch.testing.assert_close(dec_n[0].weight.grad, g2[:64, :128])
    assert id(dec_t[2].weight) not in live           # the inserted identity layer is a constant


def test_train_mode_batchnorm_statistics_round_trip():
    _, _, ours, twin = _pair(128, "GraphSage_meanAggr")
    ours.train()
    twin.sync(differentiable=False)
    bt, bn_ = twin.bn_pairs[0]
    with torch.no_grad():
        bt.running_mean[:128] += 1.0
        bt.num_batches_tracked += 1
    before = bn_.running_mean.clone()
    twin.copy_back_buffers()
    torch.testing.assert_close(bn_.running_mean, before + 1.0)
    assert int(bn_.num_batches_tracked) == 1


def test_hidden_between_129_and_255_fails_like_the_reference():
    m = BuckGNN(16, 5, 200, 2, "mean", model_name="GraphSage_meanAggr")
    assert not hasattr(m, "node_encoder")            # Models/BuckGNN.py:41,67: no branch builds one
