"""Guard-band runs: the bounds check of this repo's kernels.

compute-sanitizer is closed on the GPU pool (tools/gpu_sanitize.sh is refused there), so the out-of-bounds and
uninitialised-read checks SURVEY.md section 5 asks for are done with our own instrumentation: while a forward or a
training step runs, every CUDA tensor the package allocates (`torch.empty` / `torch.zeros` / `torch.empty_like`) is
carved out of a larger buffer with a 4 KB canary band directly before its first and directly after its last byte
(no alignment slack behind the tensor: a store one element past the end hits the band), and `empty` bodies are
pre-filled with NaN (float types), so
  * a kernel that STORES outside one of its tensors (tail tiles, ragged batches, the last band of a persistent
    kernel, workspace slots) trips a canary,
  * a kernel that READS a float it was supposed to have been given by an earlier kernel, but was not, turns the
    prediction into NaN or breaks the parity with the oracle,
and both are asserted after the run: the canaries are intact, and the instrumented run's result is BIT-IDENTICAL to a
plain run of the same call (the kernels are deterministic, so any dependence on what an `empty` tensor held before
shows up here) and agrees with the oracle.  (Under torch's caching allocator either bug would silently touch a
recycled neighbour block.)  The oracle bar in this file is a loose 1e-2: the ragged batch holds 5- and 7-node graphs,
on which a 16-bit mode's rounding does not average out (profiles/r02_precision_probe_ragged.txt); the 1e-3 / 1e-4
parity tests proper are tests/test_gpu_forward.py and friends."""
import contextlib

import pytest
import torch

from buckgnn_b200.model import BuckGNN
from buckgnn_b200.synth import collate, make_batch, make_plate_graph
from oracle.buckgnn_oracle import OracleBuckGNN, randomize_bn_stats

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PAD = 4096
CANARY = 0xA5


class _Guard:
    def __init__(self):
        self.records = []          # (buffer, nbytes)
        self.n_float_empty = 0

    def alloc(self, shape, dtype, device, zero, orig_empty):
        if isinstance(shape, int):
            shape = (shape,)
        shape = tuple(int(s) for s in shape)
        numel = 1
        for s in shape:
            numel *= s
        itemsize = torch.empty(0, dtype=dtype).element_size()
        nbytes = numel * itemsize
        buf = orig_empty(nbytes + 2 * PAD, dtype=torch.uint8, device=device)
        buf[:PAD] = CANARY
        buf[PAD + nbytes:] = CANARY
        body = buf[PAD:PAD + nbytes]
        if zero:
            body.zero_()
        elif dtype.is_floating_point:
            body.fill_(0xFF)           # fp16 / bf16 / fp32 / fp64 NaN
            self.n_float_empty += 1
        else:
            body.zero_()               # (an uninitialised INDEX would be dereferenced: keep those runs alive)
        self.records.append((buf, nbytes))
        return body.view(dtype).view(shape)

    def check(self):
        torch.cuda.synchronize()
        bad = []
        for i, (buf, nbytes) in enumerate(self.records):
            head_ok = bool((buf[:PAD] == CANARY).all())
            tail_ok = bool((buf[PAD + nbytes:] == CANARY).all())
            if not (head_ok and tail_ok):
                where = []
                if not head_ok:
                    where.append(f"before (last clobbered byte at -{PAD - int((buf[:PAD] != CANARY).nonzero().min())})")
                if not tail_ok:
                    where.append(f"after (first clobbered byte at +{int((buf[PAD + nbytes:] != CANARY).nonzero().min())})")
                bad.append(f"allocation #{i} of {nbytes} bytes: write {' and '.join(where)}")
        assert not bad, "out-of-bounds device writes:\n" + "\n".join(bad)
        assert len(self.records) > 0


@contextlib.contextmanager
def guarded():
    g = _Guard()
    orig_empty, orig_zeros, orig_empty_like = torch.empty, torch.zeros, torch.empty_like

    def is_cuda(device):
        return device is not None and torch.device(device).type == "cuda"

    def shape_of(args):
        return args[0] if len(args) == 1 and not isinstance(args[0], int) else args

    def make(zero, orig):
        def f(*args, **kw):
            dev = kw.get("device")
            if not is_cuda(dev) or kw.get("pin_memory") or kw.get("out") is not None or not args:
                return orig(*args, **kw)
            extra = set(kw) - {"device", "dtype", "requires_grad"}
            if extra:
                return orig(*args, **kw)
            t = g.alloc(shape_of(args), kw.get("dtype") or torch.get_default_dtype(), dev, zero, orig_empty)
            return t.requires_grad_() if kw.get("requires_grad") else t
        return f

    def empty_like(t, **kw):
        dev = kw.get("device", t.device)
        if not is_cuda(dev) or (set(kw) - {"device", "dtype"}) or not t.is_contiguous():
            return orig_empty_like(t, **kw)
        return g.alloc(tuple(t.shape), kw.get("dtype") or t.dtype, dev, False, orig_empty)

    torch.empty, torch.zeros, torch.empty_like = make(False, orig_empty), make(True, orig_zeros), empty_like
    try:
        yield g
    finally:
        torch.empty, torch.zeros, torch.empty_like = orig_empty, orig_zeros, orig_empty_like


def _pair(model_name, precision, layers=3, hidden=512, pooling="mean", seed=0, cls_kw=None, **kw):
    torch.manual_seed(seed)
    cfg = dict(num_node_features=16, num_edge_features=5, hidden_channels=hidden, num_layers=layers,
               pooling_layer=pooling, model_name=model_name, **(cls_kw or {}))
    ref = OracleBuckGNN(**cfg).eval()
    randomize_bn_stats(ref, realistic=True)
    ours = BuckGNN(**cfg, precision=precision, **kw)
    ours.load_state_dict(ref.state_dict())
    return ref, ours.to(DEV).eval()


def _ragged(stiffened=False):
    """graph sizes that leave partial 32-row blocks, partial 128 / 256-row tiles and a last persistent band that is
    shorter than the others"""
    sizes = [(3, 2), (31, 17), (2, 2), (40, 33), (9, 7), (23, 29)]
    return collate([make_plate_graph(i, nx=nx, ny=ny, stiffened=stiffened) for i, (nx, ny) in enumerate(sizes)])


def _rel(got, want):
    return ((got - want).abs() / want.abs().clamp(min=1e-3)).max().item()


def _guarded_forward(ref, ours, b, batch_none=False):
    """(result of the instrumented run, oracle result); asserts canaries, finiteness and bit-identity with a plain run"""
    with torch.no_grad():
        want, _ = ref(b.x, b.edge_index, b.edge_attr, None if batch_none else b.batch)
        bd = b.to(DEV)
        bvec = None if batch_none else bd.batch
        plain, _ = ours(bd.x, bd.edge_index, bd.edge_attr, bvec)
        plain = plain.float().cpu()
        with guarded() as g:
            got, _ = ours(bd.x, bd.edge_index, bd.edge_attr, bvec)
            got = got.float().cpu()
            g.check()
    assert g.n_float_empty > 0                      # the instrumentation did see the engine's activations
    assert torch.isfinite(got).all(), "a kernel read floats nobody wrote"
    assert torch.equal(got, plain), "the result depends on what an uninitialised tensor held"
    return got, want


@pytest.mark.parametrize("name,precision,rtol", [
    ("GraphSage_meanAggr", "fp16", 1e-3), ("GraphSage_meanAggr", "bf16", 3e-3), ("GraphSage_meanAggr", "tf32", 1e-3),
    ("GraphSage_meanAggr", "fp32", 1e-4), ("GraphSage_maxAggr", "fp16", 1e-3), ("GraphSage_sumAggr", "tf32", 1e-3),
    ("GraphSage_addAggr_Shared", "tf32", 1e-3), ("EA_GNN", "fp16", 1e-3), ("EA_GNN_Shared", "tf32", 1e-3),
    ("GraphSAGE_SAG", "fp32", 1e-4), ("EAGNN_SAG", "tf32", 1e-3)])
def test_forward_stays_inside_its_tensors(name, precision, rtol):
    ref, ours = _pair(name, precision)
    got, want = _guarded_forward(ref, ours, _ragged(stiffened=name.startswith("EA")))
    assert _rel(got, want) < max(rtol, 1e-2 if precision != "fp32" else rtol)


@pytest.mark.parametrize("pooling", ["supernode_only", "supernode_with_pooling", "mean_no_super", "mlp", "mlp_no_super"])
def test_pooling_variants_stay_inside_their_tensors(pooling):
    ref, ours = _pair("GraphSage_meanAggr", "fp16", pooling=pooling)
    got, want = _guarded_forward(ref, ours, _ragged())
    assert _rel(got, want) < 1e-2


@pytest.mark.parametrize("hidden", [128, 256])
def test_narrow_hidden_stays_inside_its_tensors(hidden):
    ref, ours = _pair("GraphSage_meanAggr", "fp16", hidden=hidden)
    got, want = _guarded_forward(ref, ours, _ragged())
    assert _rel(got, want) < 1e-2


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_fused_sage_layer_stays_inside_its_tensors(precision):
    """bg_sage_fused512 + bg_sage_aggregate_hubs (model.fuse_aggregate): side buffer, workspace and output rows"""
    ref, ours = _pair("GraphSage_meanAggr", precision, layers=4)
    ours.fuse_aggregate = True
    got, want = _guarded_forward(ref, ours, _ragged())
    assert _rel(got, want) < 1e-2


def test_single_tiny_graph_and_batch_none():
    ref, ours = _pair("GraphSage_meanAggr", "fp16")
    got, want = _guarded_forward(ref, ours, make_batch(1, nx=2, ny=2), batch_none=True)
    assert got.dim() == 0 and _rel(got, want) < 1e-2


@pytest.mark.parametrize("name,precision", [("GraphSage_meanAggr", "tf32"), ("GraphSage_meanAggr", "bf16"),
                                            ("GraphSage_maxAggr", "tf32"), ("EA_GNN", "tf32")])
def test_training_step_stays_inside_its_tensors(name, precision):
    cfg = dict(num_node_features=16, num_edge_features=5, hidden_channels=512, num_layers=3,
               pooling_layer="mean", model_name=name, dropout_rate=0.1)
    bd = _ragged(stiffened=name.startswith("EA")).to(DEV)
    y = torch.rand(bd.num_graphs, generator=torch.Generator().manual_seed(5)).to(DEV) + 0.5

    def two_steps(instrumented):
        torch.manual_seed(1)
        torch.cuda.manual_seed_all(1)
        ours = BuckGNN(**cfg, train_precision=precision).to(DEV).train()
        opt = torch.optim.Adam(ours.parameters(), lr=1e-3)
        with (guarded() if instrumented else contextlib.nullcontext()) as g:
            for _ in range(2):
                opt.zero_grad(set_to_none=True)
                pred, _ = ours(bd.x, bd.edge_index, bd.edge_attr, bd.batch)
                loss = ((pred - y) ** 2).mean()
                loss.backward()
                opt.step()
            if instrumented:
                g.check()
        return ours, loss

    plain, _ = two_steps(False)
    ours, loss = two_steps(True)
    assert torch.isfinite(loss).item()
    for (n_, p), (_, q) in zip(ours.named_parameters(), plain.named_parameters()):
        assert torch.isfinite(p).all(), f"non-finite parameter {n_}"
        assert torch.equal(p, q), f"{n_}: the training step depends on what an uninitialised tensor held"
    for (n_, p), (_, q) in zip(ours.named_buffers(), plain.named_buffers()):
        assert torch.equal(p, q), f"buffer {n_}"


def test_device_collate_and_wire_expansion_stay_inside_their_tensors():
    from buckgnn_b200.collate import DeviceGraphStore
    from buckgnn_b200.pipeline import WireBatch
    graphs = [make_plate_graph(i, nx=4 + i, ny=3 + (i % 3)) for i in range(7)]
    with guarded() as g:
        store = DeviceGraphStore(graphs, DEV)
        sel = torch.tensor([6, 0, 3, 3, 1])
        got = store.batch(sel)
        g.check()
    want = collate([graphs[int(i)] for i in sel])
    assert torch.equal(got.edge_index.cpu(), want.edge_index) and torch.equal(got.batch.cpu(), want.batch)
    w = WireBatch.from_batch(want)
    dev_w = WireBatch(*[getattr(w, f).to(DEV) for f in WireBatch.FIELDS], w.num_graphs, w.num_nodes, w.num_edges, w.edge_features)
    with guarded() as g:
        full = dev_w.expand(DEV)
        g.check()
    assert torch.equal(full.edge_index.cpu(), want.edge_index) and torch.equal(full.batch.cpu(), want.batch)


def test_the_guard_itself_catches_a_store_past_the_end():
    """Self-test of the instrumentation: a C-ABI call told to convert 4 elements more than the destination holds."""
    from buckgnn_b200 import capi
    src = torch.randn(260, device=DEV)
    with guarded() as g:
        dst = torch.empty(256, dtype=torch.float16, device=DEV)
        assert torch.isnan(dst).all()                                  # `empty` bodies are NaN-filled
        capi.cast_f32(src.data_ptr(), dst.data_ptr(), capi.BG_F16, 256, torch.cuda.current_stream().cuda_stream)
        g.check()                                                      # in bounds: passes
        assert torch.equal(dst, src[:256].half())
        capi.cast_f32(src.data_ptr(), dst.data_ptr(), capi.BG_F16, 260, torch.cuda.current_stream().cuda_stream)
        with pytest.raises(AssertionError, match=r"after \(first clobbered byte at \+0\)"):
            g.check()
