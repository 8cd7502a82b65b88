"""buckgnn_b200: B200-native (sm_100a) implementation of the BuckGNN forward hot path.

Layout:
  csrc/      hand-written CUDA kernels + the C ABI (include/buckgnn_b200.h)
  capi.py    ctypes binding of the C ABI (device pointers in, status codes out)
  engine.py  the forward pass as a sequence of C-ABI calls on the current CUDA stream
  model.py   `BuckGNN` nn.Module with the reference's constructor / forward / state_dict
  synth.py   synthetic plate-mesh batches in the reference's data layout
"""
__version__ = "0.1.0"
