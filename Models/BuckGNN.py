"""`from Models.BuckGNN import BuckGNN` (reference `INFERENCE.py:71`) -> the sm_100a build."""
from buckgnn_b200.model import BuckGNN, GraphNetBlock, MLPPooling, SAGEConv  # noqa: F401
