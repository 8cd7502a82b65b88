"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the
header declares; the module keeps the reference's contract; nothing falls back to CPU."""
import ctypes
import os
import re

import pytest
import torch

import __graft_entry__ as entry
from buckgnn_b200 import capi
from buckgnn_b200.model import BuckGNN
from buckgnn_b200.synth import make_batch
from oracle.buckgnn_oracle import OracleBuckGNN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    entry.build()
    return capi.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "buckgnn_b200.h")).read()
    declared = set(re.findall(r"\b(bg_[a-z0-9_]+)\s*\(", header))
    assert declared == set(capi.EXPORTED_SYMBOLS)
    raw = ctypes.CDLL(capi.library_path())
    for name in declared:
        assert hasattr(raw, name), name
    assert lib.bg_abi_version() == capi.ABI_VERSION == 11


def test_size_queries_and_argument_errors_without_gpu(lib):
    assert capi.csr_max_big_rows(650) == 11
    assert capi.csr_workspace_bytes(1000, 5000) >= 2 * 4 * 1001
    assert capi.aggregate_workspace_bytes(3) >= 3 * 16 * 512 * 4
    assert capi.pool_workspace_bytes(4) >= 4 * 8 * 512 * 4
    with pytest.raises(capi.BuckGNNError) as e:
        capi.csr_workspace_bytes(-1, 0)
    assert e.value.status == -1 and "bad argument" in str(e.value)
    # bad enum / null pointers are rejected before any CUDA call
    with pytest.raises(capi.BuckGNNError):
        capi.gemm512([], 128, capi.BG_BF16, capi.BG_BF16, 0, capi.BG_BF16, 512, None)
    with pytest.raises(capi.BuckGNNError):
        capi.csr_build(None, 10, 10, 7, None, None, None, None, None, None, 0, None)


def test_training_and_collate_entry_points_validate_arguments_without_gpu(lib):
    """Every argument check below happens before the first CUDA call (status -1 / -3 / -4, never a crash)."""
    assert capi.train_workspace_bytes(1000) > 0
    assert capi.colsum_workspace_bytes(1000, 512) >= 512 * 4
    assert capi.sgemm_workspace_bytes(64, 64, 1000) >= 64 * 64 * 4
    assert capi.hubfold_workspace_bytes(100000, capi.BG_F16, 16, 6000) >= capi.aggregate_workspace_bytes(16)
    bad = [
        lambda: capi.bn_batch_stats(None, capi.BG_F32, 0, None, None, 1e-5, 0.1, None, None, None, None, None, None, None, None, 0, None),
        lambda: capi.bn_act_forward(None, None, None, capi.BG_F32, 10, None, None, 1.5, 0, None),          # dropout_p >= 1
        lambda: capi.sage_backward_rows(None, None, None, None, None, capi.BG_F32, 0, None, None, None, None, 0.0, 0,
                                        None, None, False, None, None, None, None, 0, None),
        lambda: capi.transpose_chunks(1, capi.BG_F32, 100, 500, 500, 1, 128, 1, None),                     # n_cols % 32
        lambda: capi.transpose_chunks(1, capi.BG_F32, 100, 512, 512, 1, 64, 1, None),                      # chunks too short
        lambda: capi.wgrad512(16, 512, 16, 512, 512, capi.BG_F16, 100, 1, 100, 16, None),                  # chunk_k % 64
        lambda: capi.wgrad512(16, 512, 16, 513, 513, capi.BG_F16, 100, 1, 128, 16, None),                  # act_cols > 512
        lambda: capi.pool_backward(None, 512, None, 0, 0, 10, None, capi.BG_F32, None),                    # G = 0
        lambda: capi.pool_backward(1, 512, 1, 4, 3, 10, 16, capi.BG_F32, None),                            # concat pooling needs ldp 1024
        lambda: capi.sgemm(None, 7, 1, 1, None, capi.BG_F32, 1, 1, 4, 4, 4, None, False, None, capi.BG_F32, 0, 16,
                           capi.BG_F32, 4, False, None, 0, None),                                          # bad dtype
        lambda: capi.eigen_loss(None, None, 0, 1.0, 0.0, 1e-8, None, None, None, None),
        lambda: capi.collate_ptr(None, 3, None, None, None, None, None),
        lambda: capi.dropout_mask(0, 1.0, 10, 16, None),
        lambda: capi.mask_narrow(1, capi.BG_F32, 64, None, capi.BG_F32, 0, 10, 128, 16, None),             # ld_in < n_cols
        lambda: capi.publish_words(None, None, 8, None),
        lambda: capi.sage_aggregate(16, 16, capi.BG_F16, 10, 16, 16, 16, 0, capi.BG_AGGR_MEAN, 16, 0, None, width=100),
        # SAGPooling entry points
        lambda: capi.sag_workspace_bytes(-1, 0, 0),
        lambda: capi.sag_select(16, capi.BG_F32, 10, 16, 16, 16, 0, 16, 16, 0.0, 1.0, 16, 1, 1.5, 16, 4, 16, 16, 16, 16, 16, 16, 16, 16,
                                1 << 20, None),                                                            # ratio > 1
        lambda: capi.sag_select(16, capi.BG_F32, 10, 16, 16, 16, 0, 16, 16, 0.0, 1.0, 16, 1, 0.5, 16, 4, 16, 16, 16, 16, 16, 16, 16, 16,
                                8, None),                                                                  # workspace too small
        lambda: capi.sag_connect(16, 4, 10, 16, 5, 16, None, 16, 1 << 20, None),                           # E' > E
        lambda: capi.gather_rows(16, capi.BG_F16, 100, 16, None, 4, 16, 512, None),                        # ldx < 512
        lambda: capi.gather_rows(None, capi.BG_F16, 512, 16, None, 4, 16, 512, None),
        lambda: capi.index_invert(None, 4, None, None),
        lambda: capi.index_gather(None, None, 4, None, None),
        lambda: capi.max_aggregate_backward(None, 16, 16, capi.BG_F32, 10, 16, 16, None, 0, 16, 16, None, 0, 16, 16, None, 0, None),
        lambda: capi.max_aggregate_backward(16, 16, 16, capi.BG_F32, 10, 16, 16, None, 3, 16, 16, None, 0, 16, 16, 16, 1 << 20, None),   # big rows missing
        lambda: capi.max_aggregate_backward(16, 16, 16, capi.BG_F32, 10, 16, 16, 16, 3, 16, 16, None, 0, 16, 16, 16, 8, None),           # workspace too small
        lambda: capi.max_bwd_workspace_bytes(-1, 0),
        lambda: capi.sag_pool_backward(16, 16, capi.BG_F32, 10, 11, 16, 16, 16, 1.0, 16, 16, None, 0, 16, 16, 16, 16, 16, None),   # N' > N
        lambda: capi.sag_pool_backward(None, None, capi.BG_F32, 10, 5, 16, 16, 16, 1.0, 16, 16, None, 0, 16, 16, 16, 16, 16, None),
    ]
    for i, call in enumerate(bad):
        with pytest.raises(capi.BuckGNNError) as e:
            call()
        assert e.value.status in (-1, -3, -4), (i, e.value.status)


@pytest.mark.parametrize("name", ["GraphSage_meanAggr", "GraphSage_sumAggr", "GraphSage_addAggr",
                                  "GraphSage_maxAggr", "GraphSage_addAggr_Shared", "EA_GNN", "EA_GNN_Shared",
                                  "GraphSAGE_MLP", "GraphSAGE_SAG", "EAGNN_SAG"])
@pytest.mark.parametrize("hidden", [128, 512])
def test_state_dict_layout_matches_oracle_restatement(name, hidden):
    kw = dict(num_node_features=16, num_edge_features=5, hidden_channels=hidden, num_layers=3,
              pooling_layer="mean", model_name=name)
    ours, ref = BuckGNN(**kw), OracleBuckGNN(**kw)
    a, b = ours.state_dict(), ref.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    ours.load_state_dict(b, strict=True)
    assert all(torch.equal(ours.state_dict()[k], b[k]) for k in b)


def test_reference_import_paths_and_positional_ctor():
    from Models.BuckGNN import BuckGNN as A
    from Models.EA_GNN import EdgeAugmentedGNN as B
    assert A is B is BuckGNN
    # INFERENCE.py:73-84 / TRAIN_FINAL.py:161-166 call order
    m = A(16, 5, 512, 6, "mean", prediction_type="buckling", use_z_coord=False, use_rotations=False,
          dropout_rate=0.1, model_name="GraphSage_meanAggr")
    assert m.hidden_channels == 512 and m.num_layers == 6 and m.dropout.p == 0.1
    assert [p.requires_grad for p in m.parameters()].count(False) == 0


def test_cpu_tensors_fail_loudly_no_fallback():
    b = make_batch(1, nx=4, ny=4)
    m = BuckGNN(16, 5, 512, 2, "mean", model_name="GraphSage_meanAggr").eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        m(b.x, b.edge_index, b.edge_attr, b.batch)


def test_product_code_never_imports_the_oracle():
    for d in ("buckgnn_b200", "Models"):
        for f in os.listdir(os.path.join(ROOT, d)):
            if f.endswith(".py"):
                src = open(os.path.join(ROOT, d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_sag_variants_register_the_reference_module_names():
    """reference :190-244: first-half layers, `pool` (PyG SAGPooling with a SAGEConv(h, 1) scorer), second-half layers,
    registered after `pooling_mpl`; num_layers // 2 layers come before the pooling."""
    m = BuckGNN(16, 5, 512, 5, "mean", model_name="GraphSAGE_SAG")
    keys = list(m.state_dict().keys())
    assert len(m.sage_layers_1) == 2 and len(m.sage_layers_2) == 3 and len(m.batch_norms_1) == 2 and len(m.batch_norms_2) == 3
    assert keys.index("pooling_mpl.mlp.0.bias") < keys.index("sage_layers_1.0.lin_l.weight") < keys.index("pool.gnn.lin_l.weight") \
        < keys.index("sage_layers_2.0.lin_l.weight")
    assert tuple(m.pool.gnn.lin_l.weight.shape) == (1, 512) and tuple(m.pool.gnn.lin_l.bias.shape) == (1,)
    assert "pool.gnn.lin_r.bias" not in keys and m.precision == "fp32"
    e = BuckGNN(16, 5, 512, 4, "mean", model_name="EAGNN_SAG")
    assert len(e.gnn_layers_1) == 2 and len(e.gnn_layers_2) == 2 and len(e.batch_norms_1) == 0
    assert "gnn_layers_2.1.node_mlp_beta.2.bias" in e.state_dict()
    # a PyG >= 2.4 checkpoint carries pool.select.weight: accepted with strict=True, only its sign matters
    sd = dict(m.state_dict())
    sd["pool.select.weight"] = torch.tensor([[-2.0]])
    m.load_state_dict(sd, strict=True)
    assert m._sag_sign == -1.0


@pytest.mark.parametrize("name,pooling,ptype", [
    ("GraphSage_meanAggr", "mean", "buckling"), ("GraphSage_maxAggr", "mlp", "buckling"),
    ("GraphSage_addAggr_Shared", "supernode_with_pooling", "buckling"), ("EA_GNN", "mean", "buckling"),
    ("EA_GNN_Shared", "mlp_no_super", "buckling"), ("GraphSAGE_SAG", "mean", "buckling"), ("EAGNN_SAG", "mean", "buckling"),
    ("GraphSage_sumAggr", "mean", "static_disp"), ("EA_GNN", "supernode_only", "mode_shape"),
])
def test_trainable_parameters_are_exactly_what_autograd_reaches_in_the_oracle(name, pooling, ptype):
    """`train.trainable_parameters` (the flat gradient bucket of the all-reduce, and the tensors the training step
    returns gradients for) must list exactly the parameters `loss.backward()` reaches in the reference forward --
    the reference registers more modules than a given model_name / pooling_layer uses (Models/BuckGNN.py:164,184-187)."""
    from buckgnn_b200 import train
    kw = dict(num_node_features=16, num_edge_features=5, hidden_channels=512, num_layers=4, pooling_layer=pooling,
              prediction_type=ptype, model_name=name, dropout_rate=0.0)
    torch.manual_seed(0)
    ref = OracleBuckGNN(**kw).train()
    b = make_batch(2, nx=4, ny=3)
    out, _ = ref(b.x, b.edge_index, b.edge_attr, b.batch)
    out.square().sum().backward()
    reached = {k for k, p in ref.named_parameters() if p.grad is not None}
    ours = BuckGNN(**kw)
    ids = {id(p) for p in train.trainable_parameters(ours)}
    listed = {k for k, p in ours.named_parameters() if id(p) in ids}
    assert listed == reached, (sorted(listed - reached), sorted(reached - listed))


def test_ctypes_signatures_match_the_header_prototypes():
    """every prototype of include/buckgnn_b200.h has a ctypes signature with the same number of arguments, pointer
    arguments bound as pointers and 64-bit sizes as int64 (a mismatch would corrupt the call frame silently)"""
    import ctypes as C
    header = open(os.path.join(ROOT, "include", "buckgnn_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = re.findall(r"\b(?:int|int64_t|const char\*)\s+(bg_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)
    assert len(protos) == len(capi.EXPORTED_SYMBOLS)
    for name, params in protos:
        params = " ".join(params.split())
        args = [] if params in ("", "void") else [a.strip() for a in params.split(",")]
        _, argtypes = capi._SIGNATURES[name]
        assert len(argtypes) == len(args), (name, len(argtypes), args)
        for a, t in zip(args, argtypes):
            if "*" in a or a.startswith("void*"):
                assert t in (C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.POINTER(capi.GemmSegment),
                             C.POINTER(capi.Epilogue), C.POINTER(capi.FusedAggregate)), (name, a, t)
            elif a.startswith("int64_t"):
                assert t is C.c_int64, (name, a, t)
            elif a.startswith("int32_t"):
                assert t is C.c_int32, (name, a, t)
            elif a.startswith("size_t"):
                assert t is C.c_size_t, (name, a, t)
            elif a.startswith("uint64_t"):
                assert t is C.c_uint64, (name, a, t)
            elif a.startswith("float"):
                assert t is C.c_float, (name, a, t)
            elif a.startswith("int "):
                assert t is C.c_int, (name, a, t)
            else:
                raise AssertionError(f"unclassified parameter {a!r} of {name}")


def test_gemm_register_budget_by_warpgroup():
    """setmaxnreg acts on whole warpgroups (4 warps) and `inc` BLOCKS until the CTA's register pool has room: a budget
    above 64 K registers, or two different values inside one warpgroup, hangs the kernel with no watchdog able to
    see it (round 2: the first fused-layer build).  Static check of the numbers in gemm_tc.cuh."""
    import re
    src = open(os.path.join(ROOT, "buckgnn_b200", "csrc", "gemm_tc.cuh")).read()
    dec_producer = int(re.search(r"else if \(warp < kEpiFirstWarp\) \{\s*setmaxnreg_dec<(\d+)>", src).group(1))
    dec_gather = int(re.search(r"if constexpr \(kFuse\) \{\s*setmaxnreg_dec<(\d+)>", src).group(1))
    inc_fused, inc_plain = map(int, re.search(r"if constexpr \(kFuse\) setmaxnreg_inc<(\d+)>\(\); else setmaxnreg_inc<(\d+)>\(\);", src).groups())
    threads_plain = int(re.search(r"constexpr int kGemmThreads = (\d+);", src).group(1))
    threads_fused = int(re.search(r"constexpr int kGemmThreadsFused = (\d+);", src).group(1))
    gather_warps = int(re.search(r"constexpr int kGatherWarps = (\d+);", src).group(1))
    for v in (dec_producer, dec_gather, inc_fused, inc_plain):
        assert 24 <= v <= 256 and v % 8 == 0
    assert gather_warps == 4                                               # exactly one warpgroup
    # plain kernel: warpgroup 0 (producer, MMA issuer, two idle warps) + two epilogue warpgroups
    assert threads_plain == 3 * 128 and 128 * (dec_producer + 2 * inc_plain) <= 65536
    # fused kernel: + the gather warpgroup; launch allocation = 65536 / 512 registers per thread
    assert threads_fused == 4 * 128 and 128 * (dec_producer + 2 * inc_fused + dec_gather) <= 65536
    assert dec_producer <= 65536 // threads_fused and dec_gather <= 65536 // threads_fused <= inc_fused
