"""Parity of each sm_100a kernel against the oracle, through the C ABI (needs a B200)."""
import pytest
import torch
import torch.nn.functional as F

from buckgnn_b200 import capi, engine
from buckgnn_b200.engine import Activation, build_graph_index
from buckgnn_b200.synth import make_batch
from oracle import buckgnn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _stream():
    return torch.cuda.current_stream().cuda_stream


# ----------------------------------------------------------------------------- K1
def _csr_reference(edge_index, n, key_row):
    key, other = edge_index[key_row], edge_index[1 - key_row]
    perm = torch.sort(key, stable=True).indices
    rowptr = torch.zeros(n + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(key, minlength=n), 0)
    return rowptr.int(), other[perm].int(), perm.int()


def _random_multigraph(n, e, seed, hub=None):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g)
    if hub is not None:                      # one row with a huge degree, like the super node
        ei[1, ::3] = hub
    return ei


@pytest.mark.parametrize("key_row", [1, 0])
@pytest.mark.parametrize("case", ["mesh", "random", "hub", "tiny", "empty_edges", "bighub", "hugehub", "sparsehub", "sparsehugehub"])
def test_csr_build_bit_exact(case, key_row):
    if case == "mesh":
        b = make_batch(5, nx=17, ny=13); ei, n = b.edge_index, b.num_nodes
    elif case == "random":
        n = 5000; ei = _random_multigraph(n, 60000, 1)
    elif case == "hub":
        n = 3000; ei = _random_multigraph(n, 40000, 2, hub=1234)
    elif case == "tiny":
        n = 3; ei = torch.tensor([[0, 1, 2, 2], [2, 2, 0, 1]])
    elif case == "empty_edges":
        n = 10; ei = torch.zeros((2, 0), dtype=torch.int64)
    elif case == "bighub":                    # ~50 k: not a power of two, just below the shared-memory sort capacity (57344)
        n = 200; ei = _random_multigraph(n, 150000, 3, hub=7)
    elif case == "hugehub":                   # ~67 k entries within 200 k edge ids: still the bitmap sort
        n = 200; ei = _random_multigraph(n, 200000, 4, hub=9)
    elif case == "sparsehub":                 # hub entries spread over 2.4 M edge ids (> 32 * 57 k bits): bitonic network in smem
        n = 4000; ei = _random_multigraph(n, 2400000, 5)
        ei[1, ::64] = 77
    else:                                     # ... and more entries than shared memory holds: the global-memory network
        n = 4000; ei = _random_multigraph(n, 2400000, 6)
        ei[1, ::30] = 78
    idx = build_graph_index(ei.to(DEV), None, n, key_row=key_row)
    rowptr, col, perm = _csr_reference(ei, n, key_row)
    e = ei.shape[1]
    assert torch.equal(idx.rowptr.cpu(), rowptr)
    assert torch.equal(idx.perm.cpu()[:e], perm)
    assert torch.equal(idx.col.cpu()[:e], col)
    deg = rowptr[1:] - rowptr[:-1]
    big = sorted(idx.big_rows.cpu()[:idx.n_big].tolist())
    assert big == torch.nonzero(deg > capi.BG_BIG_ROW_THRESHOLD).flatten().tolist()


def test_csr_range_hubs():
    """Super-node rows are recognised as range hubs (neighbours = one contiguous run of rows); a hub with
    arbitrary neighbours switches the whole index back to the generic hub path."""
    b = make_batch(4, nx=13, ny=11)
    idx = build_graph_index(b.edge_index.to(DEV), b.batch.to(DEV), b.num_nodes)
    assert idx.n_big == 4 and idx.hub_lo is not None and idx.hub_max_degree == 13 * 11
    big = idx.big_rows.cpu()[:4].tolist()
    lo = idx.hub_lo.cpu()[:4].tolist()
    want = torch.full((b.num_nodes,), -1, dtype=torch.int32)
    for slot, (r, l) in enumerate(zip(big, lo)):
        g = int(b.batch[r])
        assert r == int(b.ptr[g + 1]) - 1 and l == int(b.ptr[g])          # super node = last node of its graph
        want[l:r] = slot
    assert torch.equal(idx.hub_of_row.cpu()[:b.num_nodes], want)
    ei = _random_multigraph(3000, 40000, 2, hub=1234)
    idx = build_graph_index(ei.to(DEV), None, 3000)
    assert idx.n_big >= 1 and idx.hub_lo is None
    # two hubs over the same range overlap -> not foldable either
    n = 200
    src = torch.arange(100)
    ei = torch.cat([torch.stack([src, torch.full_like(src, 150)]), torch.stack([src, torch.full_like(src, 151)])], 1)
    idx = build_graph_index(ei.to(DEV), None, n)
    assert idx.n_big == 2 and idx.hub_lo is None


def test_csr_build_rejects_out_of_range_ids():
    ei = torch.tensor([[0, 1, 5], [1, 0, 1]]).to(DEV)
    with pytest.raises(IndexError):
        build_graph_index(ei, None, 3)


def test_graph_ptr_bit_exact_and_unsorted_batch_rejected():
    b = make_batch(7, nx=9, ny=8)
    idx = build_graph_index(b.edge_index.to(DEV), b.batch.to(DEV), b.num_nodes)
    assert idx.n_graphs == 7 and torch.equal(idx.graph_ptr.cpu().long(), b.ptr)
    # a graph id that never occurs (empty graph in the middle) keeps offsets monotone
    batch = torch.tensor([0, 0, 2, 2, 2, 3])
    ei = torch.zeros((2, 0), dtype=torch.int64)
    idx = build_graph_index(ei.to(DEV), batch.to(DEV), 6)
    assert idx.graph_ptr.cpu().tolist() == [0, 2, 2, 5, 6]
    with pytest.raises(ValueError):
        build_graph_index(ei.to(DEV), torch.tensor([0, 1, 0, 1, 1, 1]).to(DEV), 6)


# ----------------------------------------------------------------------------- K5 front
@pytest.mark.parametrize("n,f", [(1, 16), (63, 16), (1000, 16), (4097, 7), (777, 5), (300, 20), (50, 32)])
def test_encoder_front_matches_fp32(n, f):
    torch.manual_seed(0)
    x = torch.randn(n, f)
    l1, l2 = torch.nn.Linear(f, 64), torch.nn.Linear(64, 128)
    want = torch.relu(l2(torch.relu(l1(x)))).detach()
    out = torch.empty(n, 128, device=DEV)
    d = lambda t: t.detach().to(DEV).contiguous()
    w1, b1, w2, b2, xd = d(l1.weight), d(l1.bias), d(l2.weight), d(l2.bias), d(x)
    capi.encoder_front(xd.data_ptr(), n, f, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                       out.data_ptr(), capi.BG_F32, _stream())
    torch.testing.assert_close(out.cpu(), want, rtol=1e-5, atol=1e-5)
    for code, dt, rtol in ((capi.BG_BF16, torch.bfloat16, 8e-3), (capi.BG_F16, torch.float16, 1e-3)):
        outb = torch.empty(n, 128, device=DEV, dtype=dt)
        capi.encoder_front(xd.data_ptr(), n, f, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                           outb.data_ptr(), code, _stream())
        # 16-bit outputs come from the mma.sync kernel: x, W1, h1, W2 rounded to fp16 operands (fp32 accumulate)
        torch.testing.assert_close(outb.cpu().float(), want, rtol=rtol, atol=2e-3)
        assert ((outb.cpu().float() - want).norm() / want.norm()).item() < (4e-3 if dt == torch.bfloat16 else 1e-3)


# ----------------------------------------------------------------------------- K2
@pytest.mark.parametrize("fold", [True, False])
@pytest.mark.parametrize("precision", ["bf16", "fp16", "tf32"])
@pytest.mark.parametrize("aggr", ["mean", "sum", "max"])
def test_aggregate_matches_oracle(aggr, precision, fold):
    torch.manual_seed(1)
    b = make_batch(3, nx=21, ny=17)            # hubs of degree 357 -> the split path / the folded range-hub path
    n = b.num_nodes
    ei = torch.cat([b.edge_index, torch.tensor([[5, 5], [9, 9]])], 1)     # duplicate edge
    ei = ei[:, ei[1] != 3]                                                  # node 3 isolated
    x = torch.randn(n, 512)
    x = x.to(engine._TORCH[engine.PRECISION_FORMATS[precision][0]]).float()
    want = O.aggregate(x.double(), ei, aggr).float()
    idx = build_graph_index(ei.to(DEV), None, n)
    assert idx.n_big == 3 and idx.hub_lo is not None
    xa, oa = Activation(n, 512, precision, DEV), Activation(n, 512, precision, DEV)
    xa.data.copy_(x)
    engine.aggregate(xa, oa, idx, aggr, fold_hubs=fold)
    got = oa.data.float().cpu()
    if fold:                                    # deterministic: a second launch is bit-identical
        ob = Activation(n, 512, precision, DEV)
        engine.aggregate(xa, ob, idx, aggr, fold_hubs=True)
        assert torch.equal(ob.data, oa.data)
    if precision == "bf16":
        torch.testing.assert_close(got, want, rtol=8e-3, atol=1e-6 if aggr != "sum" else 2e-2)
    elif precision == "fp16":
        torch.testing.assert_close(got, want, rtol=1e-3, atol=1e-6 if aggr != "sum" else 3e-3)
    else:
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    if aggr == "max":
        assert got[3].abs().sum() == 0          # isolated node -> 0, not -inf


@pytest.mark.parametrize("precision", ["fp16", "bf16", "tf32"])
@pytest.mark.parametrize("aggr", ["mean", "sum", "max"])
def test_aggregate_medium_degrees(aggr, precision):
    """rows of degree 20..64 (more neighbours than one index fetch holds, still below the hub threshold) and a few hubs
    with scattered neighbours, on a random multigraph; max over 16-bit rows is exact (packed HMNMX2 path)"""
    n, e = 600, 24000
    g = torch.Generator().manual_seed(11)
    ei = torch.randint(0, n, (2, e), generator=g)
    ei[1, ::50] = 17                                       # one hub of ~480 + its random share
    x = torch.randn(n, 512, generator=g)
    x = x.to(engine._TORCH[engine.PRECISION_FORMATS[precision][0]]).float()
    want = O.aggregate(x.double(), ei, aggr).float()
    idx = build_graph_index(ei.to(DEV), None, n)
    deg = idx.rowptr.cpu()[1:] - idx.rowptr.cpu()[:-1]
    assert int(((deg > 32) & (deg <= capi.BG_BIG_ROW_THRESHOLD)).sum()) > 100 and idx.n_big >= 1
    xa, oa = Activation(n, 512, precision, DEV), Activation(n, 512, precision, DEV)
    xa.data.copy_(x)
    engine.aggregate(xa, oa, idx, aggr)
    got = oa.data.float().cpu()
    if aggr == "max":
        assert torch.equal(got, want)                      # the maximum of representable values is representable
    elif precision == "tf32":
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-4)
    else:
        torch.testing.assert_close(got, want, rtol=8e-3 if precision == "bf16" else 1e-3, atol=0.2 if aggr == "sum" else 2e-3)


@pytest.mark.parametrize("precision", ["fp16", "bf16", "tf32"])
@pytest.mark.parametrize("aggr", ["mean", "sum", "max"])
def test_aggregate_128_columns(aggr, precision):
    torch.manual_seed(7)
    b = make_batch(3, nx=19, ny=15)
    n = b.num_nodes
    ei = b.edge_index[:, b.edge_index[1] != 4]
    x = torch.randn(n, 128).to(engine._TORCH[engine.PRECISION_FORMATS[precision][0]]).float()
    want = O.aggregate(x.double(), ei, aggr).float()
    idx = build_graph_index(ei.to(DEV), None, n)
    xa, oa = Activation(n, 128, precision, DEV), Activation(n, 128, precision, DEV)
    xa.data.copy_(x)
    engine.aggregate(xa, oa, idx, aggr)
    got = oa.data.float().cpu()
    tol = {"bf16": dict(rtol=8e-3, atol=2e-2 if aggr == "sum" else 1e-6),
           "fp16": dict(rtol=1e-3, atol=3e-3 if aggr == "sum" else 1e-6),
           "tf32": dict(rtol=1e-5, atol=1e-5)}[precision]
    torch.testing.assert_close(got, want, **tol)


@pytest.mark.parametrize("precision", ["fp16", "tf32"])
def test_aggregate_folded_hubs_across_bands(precision):
    """Graphs larger and smaller than an SM's band of rows, mixed: hub ranges start and end inside bands."""
    torch.manual_seed(3)
    sizes = [(70, 64), (9, 9), (40, 33), (9, 8), (55, 61), (12, 9)] * 3
    from buckgnn_b200.synth import make_plate_graph, collate
    b = collate([make_plate_graph(i, nx=nx, ny=ny) for i, (nx, ny) in enumerate(sizes)])
    n = b.num_nodes
    x = torch.randn(n, 512).to(engine._TORCH[engine.PRECISION_FORMATS[precision][0]]).float()
    want = O.aggregate(x.double(), b.edge_index, "mean").float()
    idx = build_graph_index(b.edge_index.to(DEV), b.batch.to(DEV), n)
    assert idx.hub_lo is not None and idx.n_big == len(sizes)
    xa, oa = Activation(n, 512, precision, DEV), Activation(n, 512, precision, DEV)
    xa.data.copy_(x)
    engine.aggregate(xa, oa, idx, "mean")
    tol = dict(rtol=1e-3, atol=2e-4) if precision == "fp16" else dict(rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(oa.data.float().cpu(), want, **tol)


# ----------------------------------------------------------------------------- K3
def _gemm_case(m, precision, cta_group, full_epilogue, seed=0):
    g = torch.Generator().manual_seed(seed)
    ks = [512, 512] if full_epilogue else [128]
    a_dt, b_dt = (engine._TORCH[c] for c in engine.PRECISION_FORMATS[precision])
    dt = a_dt
    As = [(torch.randn(m, k, generator=g) / k ** 0.5).to(a_dt) for k in ks]
    Bs = [(torch.randn(512, k, generator=g)).to(b_dt) for k in ks]
    bias = torch.randn(512, generator=g) * 0.1
    acc = sum(a.double() @ b.double().T for a, b in zip(As, Bs)) + bias.double()
    if full_epilogue:
        scale = torch.rand(512, generator=g) * 20 + 5
        shift = torch.randn(512, generator=g) * 0.1
        res = torch.randn(m, 512, generator=g).to(dt)
        want = torch.relu(F.normalize(acc, dim=-1) * scale.double() + shift.double()) + res.double()
    else:
        scale = shift = res = None
        want = acc
    dv = lambda t: None if t is None else t.to(DEV).contiguous()
    Ad, Bd = [dv(a) for a in As], [dv(b) for b in Bs]
    resd = dv(res)
    hv = lambda t: None if t is None else t.float().contiguous()
    biasd, scaled, shiftd = hv(bias), hv(scale), hv(shift)          # epilogue vectors are HOST pointers
    out = Activation(m, 512, precision, DEV)
    segs = [(a.data_ptr(), k, b.data_ptr(), k, k) for a, b, k in zip(Ad, Bd, ks)]
    engine.gemm512(segs, m, precision, out, cta_group=cta_group, bias=biasd.data_ptr(),
                   bn_scale=engine._p(scaled), bn_shift=engine._p(shiftd), residual=engine._p(resd), ldr=512,
                   normalize=full_epilogue, relu=full_epilogue)
    torch.cuda.synchronize()
    return out.data.float().cpu(), want.float()


_TOL16 = {"bf16": dict(rtol=1e-2, atol=2e-2), "fp16": dict(rtol=2e-3, atol=3e-3)}


@pytest.mark.parametrize("cta_group", [2])
@pytest.mark.parametrize("precision", ["bf16", "fp16", "tf32"])
@pytest.mark.parametrize("m", [128, 300, 20000])
def test_gemm_plain_bias(m, precision, cta_group):
    got, want = _gemm_case(m, precision, cta_group, full_epilogue=False)
    tol = _TOL16.get(precision, dict(rtol=2e-3, atol=2e-3))
    torch.testing.assert_close(got, want, **tol)


@pytest.mark.parametrize("cta_group", [2])
@pytest.mark.parametrize("precision", ["bf16", "fp16", "tf32"])
@pytest.mark.parametrize("m", [77, 40000])
def test_gemm_sage_epilogue(m, precision, cta_group):
    got, want = _gemm_case(m, precision, cta_group, full_epilogue=True, seed=3)
    tol = _TOL16.get(precision, dict(rtol=2e-3, atol=3e-3))
    torch.testing.assert_close(got, want, **tol)


def test_gemm_3xtf32_is_fp32_accurate():
    """hi/lo split operands (bg_split_tf32) through the tf32 kernel: ~fp32 accuracy."""
    g = torch.Generator().manual_seed(5)
    m, k = 1000, 512
    a, w = torch.randn(m, k, generator=g), torch.randn(512, k, generator=g) / k ** 0.5
    want = (a.double() @ w.double().T).float()
    act = Activation(m, k, "fp32", DEV)
    act.data.copy_(a)
    act.refresh_split()
    pack = engine.pack_linear(w.to(DEV), "fp32")
    out = Activation(m, 512, "fp32", DEV)
    engine.gemm512(engine._segments(act, pack), m, "fp32", out)
    torch.testing.assert_close(out.data.cpu(), want, rtol=2e-5, atol=2e-5)
    trunc = (a.view(torch.int32) & ~0x1FFF).view(torch.float32)
    assert act.hi is act.data                            # the tensor core truncates: the data is its own hi part
    assert torch.equal((trunc + act.lo.cpu()), a)        # the split is exact
    hi, lo = torch.empty_like(act.data), torch.empty_like(act.data)
    capi.split_tf32(act.data.data_ptr(), hi.data_ptr(), lo.data_ptr(), act.data.numel(), _stream())
    assert torch.equal(hi.cpu(), trunc) and torch.equal(lo, act.lo)


# ----------------------------------------------------------------------------- K4
@pytest.mark.parametrize("precision", ["bf16", "fp16", "tf32"])
def test_pool_head_matches_oracle(precision):
    torch.manual_seed(2)
    b = make_batch(6, nx=11, ny=9)
    n = b.num_nodes
    x = torch.randn(n, 512)
    x = x.to(engine._TORCH[engine.PRECISION_FORMATS[precision][0]]).float()
    dec = torch.nn.Sequential(torch.nn.Linear(512, 128), torch.nn.ReLU(), torch.nn.Linear(128, 64),
                              torch.nn.ReLU(), torch.nn.Linear(64, 1))
    pooled_want = O.global_mean_pool(x, b.batch)
    want = dec(pooled_want).detach()
    idx = build_graph_index(b.edge_index.to(DEV), b.batch.to(DEV), n)
    xa = Activation(n, 512, precision, DEV)
    xa.data.copy_(x)
    d = lambda t: t.detach().to(DEV).contiguous()
    packs = {"w1": d(dec[0].weight), "b1": d(dec[0].bias), "w2": d(dec[2].weight), "b2": d(dec[2].bias),
             "w3": d(dec[4].weight), "b3": d(dec[4].bias)}
    pred, pooled = engine.pool_head(xa, idx, packs, 1, want_pooled=True)
    torch.testing.assert_close(pooled.cpu(), pooled_want, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(pred.cpu(), want, rtol=1e-4, atol=1e-5)
    # super-node variants: the last node of each graph
    last = b.ptr[1:] - 1
    keep = torch.ones(n, dtype=torch.bool); keep[last] = False
    _, pooled = engine.pool_head(xa, idx, packs, 1, want_pooled=True, pooling="supernode_only")
    torch.testing.assert_close(pooled.cpu(), x[last], rtol=0, atol=0)
    _, pooled = engine.pool_head(xa, idx, packs, 1, want_pooled=True, pooling="mean_no_super")
    torch.testing.assert_close(pooled.cpu(), O.global_mean_pool(x[keep], b.batch[keep]), rtol=1e-5, atol=1e-6)


# ----------------------------------------------------------------------------- device-side collate (row f1)
def test_device_collate_is_bit_exact():
    """DeviceGraphStore.batch == the host collate (PyG Batch.from_data_list layout), for an arbitrary selection
    with repeats and out-of-order graphs."""
    from buckgnn_b200.collate import DeviceGraphStore
    from buckgnn_b200.synth import collate, make_plate_graph
    graphs = [make_plate_graph(i, nx=5 + i % 4, ny=4 + i % 3, stiffened=(i % 2 == 1)) for i in range(9)]
    store = DeviceGraphStore(graphs, DEV)
    for sel in ([0, 1, 2, 3, 4, 5, 6, 7, 8], [7, 2, 2, 5], [3]):
        want = collate([graphs[i] for i in sel])
        for idx in (torch.tensor(sel), torch.tensor(sel, device=DEV)):     # host indices (no read-back) / device indices
            got = store.batch(idx)
            assert got.num_graphs == len(sel)
            for f in ("x", "edge_index", "edge_attr", "batch", "y", "ptr"):
                assert torch.equal(getattr(got, f).cpu(), getattr(want, f)), f
    # 1500 graphs: more than one scan block of bg_collate_ptr
    sel = torch.arange(1500) % 9
    got = store.batch(sel)
    want = collate([graphs[int(i)] for i in sel])
    assert torch.equal(got.edge_index.cpu(), want.edge_index) and torch.equal(got.ptr.cpu(), want.ptr)
    assert torch.equal(got.batch.cpu(), want.batch) and torch.equal(got.x.cpu(), want.x)


# ----------------------------------------------------------------------------- wire format (pipeline.WireBatch)
@pytest.mark.parametrize("case", ["super_node", "stiffened_virtual", "no_super", "shuffled", "single_graph", "interleaved"])
def test_wire_format_expands_to_the_exact_pyg_batch(case):
    """bg_expand_wire: explicit int32 edges + implicit hub pairs + node offsets -> the int64 edge_index and batch vector
    `batch.to(device)` would have delivered, bit for bit; graphs whose trailing edges are not the reference's hub pairs
    ship every edge explicitly."""
    from buckgnn_b200.pipeline import WireBatch
    from buckgnn_b200.synth import PlateBatch
    if case == "super_node":
        b = make_batch(5, nx=9, ny=7)
    elif case == "stiffened_virtual":
        b = make_batch(3, nx=8, ny=6, stiffened=True)
    elif case == "no_super":
        b = make_batch(3, nx=8, ny=6, stiffened=True, super_node=False, virtual_edges=True)
    elif case == "single_graph":
        b = make_batch(1, nx=10, ny=4)
    elif case == "interleaved":                              # edge list NOT grouped by graph: everything ships, as is
        b = make_batch(3, nx=8, ny=6)
        order = torch.randperm(b.num_edges, generator=torch.Generator().manual_seed(1))
        b = PlateBatch(b.x, b.edge_index[:, order].contiguous(), b.edge_attr[order], b.batch, b.y, b.ptr, b.num_graphs)
    else:                                                    # edges of every graph in random order: nothing is implicit
        b = make_batch(3, nx=8, ny=6)
        g = torch.Generator().manual_seed(0)
        eg = b.batch[b.edge_index[0]]
        order = torch.cat([torch.nonzero(eg == k).flatten()[torch.randperm(int((eg == k).sum()), generator=g)] for k in range(3)])
        b = PlateBatch(b.x, b.edge_index[:, order].contiguous(), b.edge_attr[order], b.batch, b.y, b.ptr, b.num_graphs)
    w = WireBatch.from_batch(b)
    if case in ("super_node", "stiffened_virtual", "single_graph"):
        assert w.edges.shape[1] == b.num_edges - 2 * (b.num_nodes - b.num_graphs)      # hub pairs are implicit
    else:
        assert w.edges.shape[1] == b.num_edges
    dev_w = WireBatch(*[getattr(w, f).to(DEV) for f in WireBatch.FIELDS], w.num_graphs, w.num_nodes, w.num_edges, w.edge_features)
    full = dev_w.expand(DEV)
    torch.cuda.synchronize()
    assert torch.equal(full.edge_index.cpu(), b.edge_index)
    assert torch.equal(full.batch.cpu(), b.batch)
    assert full.edge_attr.shape == (0, 5) and full.x.data_ptr() == dev_w.x.data_ptr()
