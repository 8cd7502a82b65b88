"""`from Models.EA_GNN import EdgeAugmentedGNN` (reference `TRAIN_FINAL.py:14`,
`INFERENCE_TIMER.py:177`).  The reference ships no such module; both scripts build it
with BuckGNN's constructor arguments, so it is the same class."""
from buckgnn_b200.model import BuckGNN as EdgeAugmentedGNN  # noqa: F401
