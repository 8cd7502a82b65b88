"""Pool-fused last layer (bg_epilogue.pool_block_sums, bg_pool_block_flags, bg_pool_head_blocks): the epilogue of the
last SAGE update sums its output rows per 32-row block for `global_mean_pool` (Models/BuckGNN.py:274, 515) instead of
storing them.  Checked at three levels: the flags against their definition, the GEMM's block sums and kept rows against
the unfused GEMM, and whole forwards against the unfused path and the oracle -- for every pooling variant and for
graph sizes that put boundaries inside, at the edge of and several per 32-row block."""
import pytest
import torch

from buckgnn_b200 import capi, engine
from buckgnn_b200.engine import Activation
from buckgnn_b200.model import BuckGNN
from buckgnn_b200.synth import collate, make_batch, make_plate_graph
from oracle.buckgnn_oracle import OracleBuckGNN, randomize_bn_stats

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _flags_reference(ptr, n):
    keep = torch.zeros((n + 31) // 32, dtype=torch.uint8)
    for p in ptr.tolist():
        if p < n:
            keep[p // 32] = 1
        if p > 0:
            keep[(p - 1) // 32] = 1
    return keep


@pytest.mark.parametrize("sizes", [[5, 70, 64, 1, 31, 33, 200], [32, 32, 64], [1], [1000], [3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3]])
def test_block_flags_mark_first_and_last_rows(sizes):
    ptr = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int32)
    n = int(ptr[-1])
    keep = torch.full(((n + 31) // 32,), 7, dtype=torch.uint8, device=DEV)
    capi.pool_block_flags(ptr.to(DEV).data_ptr(), len(sizes), n, keep.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert torch.equal(keep.cpu(), _flags_reference(ptr, n))


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
@pytest.mark.parametrize("m", [32 * 40, 32 * 40 + 17, 255, 4097])
def test_gemm_block_sums_and_kept_rows_match_the_unfused_epilogue(precision, m):
    g = torch.Generator().manual_seed(m)
    dt = engine._TORCH[engine.PRECISION_FORMATS[precision][0]]
    a = (torch.randn(m, 512, generator=g) / 512 ** 0.5).to(dt).to(DEV)
    x = (torch.randn(m, 512, generator=g) / 512 ** 0.5).to(dt).to(DEV)
    wl, wr = (torch.randn(512, 512, generator=g).to(dt).to(DEV) for _ in range(2))
    bias, scale, shift = torch.randn(512, generator=g), torch.rand(512, generator=g) * 20 + 5, torch.randn(512, generator=g) * 0.1
    segs = [(a.data_ptr(), 512, wl.data_ptr(), 512, 512), (x.data_ptr(), 512, wr.data_ptr(), 512, 512)]
    kw = dict(bias=bias.data_ptr(), bn_scale=scale.data_ptr(), bn_shift=shift.data_ptr(), normalize=True, relu=True)
    plain = Activation(m, 512, precision, DEV)
    engine.gemm512(segs, m, precision, plain, **kw)
    nb = (m + 31) // 32
    keep = (torch.rand(nb, generator=g) < 0.3).to(torch.uint8).to(DEV)
    sums = torch.full((nb, 512), float("nan"), dtype=torch.float32, device=DEV)
    fused = Activation(m, 512, precision, DEV)
    fused.data.fill_(-7.0)
    engine.gemm512(segs, m, precision, fused, pool_block_sums=sums.data_ptr(), pool_block_keep=keep.data_ptr(), **kw)
    torch.cuda.synchronize()
    rows_kept = keep.bool().repeat_interleave(32)[:m]
    assert torch.equal(fused.data[rows_kept], plain.data[rows_kept])              # kept blocks: identical rows
    assert bool((fused.data[~rows_kept] == -7.0).all())                            # other blocks: never written
    want = torch.zeros(nb * 32, 512, dtype=torch.float64, device=DEV)
    want[:m] = plain.data.double()
    want = want.view(nb, 32, 512).sum(1)
    err = (sums.double() - want).abs().max().item()
    assert err <= 1e-5 * max(want.abs().max().item(), 1.0), err


def _pair(pooling, precision="fp16", layers=3, **kw):
    torch.manual_seed(0)
    cfg = dict(num_node_features=16, num_edge_features=5, hidden_channels=512, num_layers=layers,
               pooling_layer=pooling, model_name="GraphSage_meanAggr")
    ref = OracleBuckGNN(**cfg).eval()
    randomize_bn_stats(ref, realistic=True)
    out = []
    for fuse in (True, False):
        m = BuckGNN(**cfg, precision=precision, fuse_pool=fuse, **kw)
        m.load_state_dict(ref.state_dict())
        out.append(m.to(DEV).eval())
    return ref, out[0], out[1]


def _rel(got, want):
    return ((got - want).abs() / want.abs().clamp(min=1e-3)).max().item()


RAGGED = [(3, 2), (31, 17), (2, 2), (40, 33), (5, 6), (8, 4), (4, 8), (16, 16)]     # 6+1 .. 1320+1 nodes per graph


@pytest.mark.parametrize("pooling", ["mean", "mean_no_super", "supernode_only", "supernode_with_pooling", "mlp", "mlp_no_super"])
def test_fused_pool_forward_equals_unfused_and_oracle(pooling):
    ref, fused, unfused = _pair(pooling)
    b = collate([make_plate_graph(i, nx=nx, ny=ny) for i, (nx, ny) in enumerate(RAGGED)])
    bd = b.to(DEV)
    with torch.no_grad():
        want, _ = ref(b.x, b.edge_index, b.edge_attr, b.batch)
        a, _ = fused(bd.x, bd.edge_index, bd.edge_attr, bd.batch)
        u, _ = unfused(bd.x, bd.edge_index, bd.edge_attr, bd.batch)
    assert _rel(a.cpu(), u.cpu()) < 2e-6                      # same rows, another fp32 summation order
    assert _rel(a.cpu(), want) < 1e-3


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_fused_pool_batch_none_and_single_graph(precision):
    ref, fused, unfused = _pair("mean", precision)
    g = make_plate_graph(0, nx=21, ny=13)
    with torch.no_grad():
        a, _ = fused(g.x.to(DEV), g.edge_index.to(DEV), g.edge_attr.to(DEV), None)
        u, _ = unfused(g.x.to(DEV), g.edge_index.to(DEV), g.edge_attr.to(DEV), None)
    assert a.dim() == u.dim() == 0 and _rel(a.cpu(), u.cpu()) < 2e-6


def test_fused_pool_is_the_default_and_is_skipped_where_it_cannot_apply():
    from buckgnn_b200.engine import TIMERS
    _, fused, _ = _pair("mean")
    b = make_batch(3, nx=12, ny=10).to(DEV)
    TIMERS.enable()
    with torch.no_grad():
        fused(b.x, b.edge_index, b.edge_attr, b.batch)
    names = set(TIMERS.summary())
    TIMERS.disable()
    assert "sage_update_pool" in names
    _, tf32, _ = _pair("mean", "tf32")                          # fp32 storage: no pool-fused epilogue, plain path
    TIMERS.enable()
    with torch.no_grad():
        tf32(b.x, b.edge_index, b.edge_attr, b.batch)
    names = set(TIMERS.summary())
    TIMERS.disable()
    assert "sage_update_pool" not in names


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fused_pool_random_ragged_batches(seed):
    """random graph sizes from 2 to ~700 nodes (many boundaries per 32-row block, blocks shared by three graphs,
    boundaries on multiples of 32): fused == unfused for the mean and the super-node variants"""
    g = torch.Generator().manual_seed(seed)
    dims = [(int(a), int(b)) for a, b in torch.randint(1, 27, (40, 2), generator=g).tolist()]
    dims += [(31, 1), (1, 1), (15, 2), (3, 5)]                          # 33-, 3-, 33- and 17-node graphs (+ super node)
    b = collate([make_plate_graph(i, nx=nx + 1, ny=ny + 1) for i, (nx, ny) in enumerate(dims)]).to(DEV)
    for pooling in ("mean", "supernode_with_pooling", "mean_no_super"):
        _, fused, unfused = _pair(pooling, layers=2)
        with torch.no_grad():
            a, _ = fused(b.x, b.edge_index, b.edge_attr, b.batch)
            u, _ = unfused(b.x, b.edge_index, b.edge_attr, b.batch)
        assert a.shape == u.shape == (len(dims),)
        assert _rel(a.cpu(), u.cpu()) < 2e-6, pooling
