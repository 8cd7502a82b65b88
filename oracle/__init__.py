"""CPU restatement ("oracle") of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this package, and only as the checker or the
reported CPU baseline.  Nothing under `buckgnn_b200/` or `Models/` imports it; the
product path raises if the CUDA library is missing instead of falling back here.

PARITY UNPINNED.  The arithmetic of the reference's hot path lives in two
third-party packages that are neither vendored in `/root/reference` nor installed
here (no network): `torch_geometric` (`SAGEConv`, `global_mean_pool`) and
`torch_scatter` (`scatter_mean`), both unpinned (`README.md:64-70` of the
reference names them without versions; `lin_l`/`lin_r` naming implies PyG >= 1.6 /
2.x).  The reference ships no tests, golden vectors or checkpoints.  So this oracle
restates the published semantics of those operators at the reference's call sites
(`Models/BuckGNN.py:114-176, 274, 449, 561`) and is cross-checked only against
(i) hand-computed known-answer cases and (ii) an independent dense-adjacency
formulation (`tests/test_oracle.py`), and frozen by the golden fixtures of
`tests/golden/` (generated from this oracle by `tests/golden/make_golden.py`).
"""
