"""Does tcgen05 kind::tf32 truncate fp32 operands (ignore the low 13 mantissa bits)?  If it does, the `hi` part of
the 3xTF32 split is what the tensor core reads from the raw fp32 data anyway."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buckgnn_b200 import engine
from buckgnn_b200.engine import Activation

DEV = "cuda:0"
g = torch.Generator().manual_seed(5)
m, k = 4096, 512
a, w = torch.randn(m, k, generator=g), torch.randn(512, k, generator=g) / k ** 0.5
want = (a.double() @ w.double().T)
act = Activation(m, k, "fp32", DEV)
act.data.copy_(a)
act.refresh_split()
pack = engine.pack_linear(w.to(DEV), "fp32")
w_hi, w_lo = pack.parts
out = Activation(m, 512, "fp32", DEV)
engine.gemm512(engine._segments(act, pack), m, "fp32", out)
ref = out.data.clone()
segs = [(act.data.data_ptr(), k, w_hi.data_ptr(), k, k), (act.data.data_ptr(), k, w_lo.data_ptr(), k, k),
        (act.lo.data_ptr(), k, w_hi.data_ptr(), k, k)]
engine.gemm512(segs, m, "fp32", out)
raw_w = w.to(DEV).contiguous()
print("A raw instead of A_hi: bit-identical =", torch.equal(out.data, ref),
      " max err vs fp64:", float((out.data.double().cpu() - want).abs().max()), "(split:", float((ref.double().cpu() - want).abs().max()), ")")
segs = [(act.data.data_ptr(), k, raw_w.data_ptr(), k, k), (act.data.data_ptr(), k, w_lo.data_ptr(), k, k),
        (act.lo.data_ptr(), k, raw_w.data_ptr(), k, k)]
engine.gemm512(segs, m, "fp32", out)
print("A and W raw instead of hi parts: bit-identical =", torch.equal(out.data, ref))
