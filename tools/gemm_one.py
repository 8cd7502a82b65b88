"""Run bg_gemm512 a few times on the SAGE-update shape (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buckgnn_b200 import capi, engine
from buckgnn_b200.engine import Activation
DEV = "cuda:0"
m = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
residual = (sys.argv[2] == "res") if len(sys.argv) > 2 else True
g = torch.Generator().manual_seed(0)
ks = [512, 512]
As = [(torch.randn(m, k, generator=g) / k ** 0.5).half().to(DEV) for k in ks]
Bs = [torch.randn(512, k, generator=g).half().to(DEV) for k in ks]
bias = torch.randn(512, generator=g).float(); scale = torch.rand(512, generator=g) * 20 + 5; shift = torch.randn(512, generator=g) * .1
res = torch.randn(m, 512, generator=g).half().to(DEV) if residual else None
out = Activation(m, 512, "fp16", DEV)
segs = [(a.data_ptr(), k, b.data_ptr(), k, k) for a, b, k in zip(As, Bs, ks)]
for _ in range(3):
    engine.gemm512(segs, m, "fp16", out, bias=bias.data_ptr(), bn_scale=scale.data_ptr(), bn_shift=shift.data_ptr(),
                   residual=engine._p(res), ldr=512, normalize=True, relu=True)
torch.cuda.synchronize()
print("ok")
