"""GPU micro-benchmark of bg_gemm512 alone: time per launch for the SAGE-update and encoder
shapes, plus (with the -DBG_PROFILE build) where each warp role spends its cycles."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buckgnn_b200 import capi, engine
from buckgnn_b200.engine import Activation

DEV = "cuda:0"


def case(name, m, ks, precision, normalize, residual, iters=20, pool=False):
    dt = engine._TORCH[engine.PRECISION_FORMATS[precision][0]]
    g = torch.Generator().manual_seed(0)
    As = [(torch.randn(m, k, generator=g) / k ** 0.5).to(dt).to(DEV) for k in ks]
    Bs = [torch.randn(512, k, generator=g).to(dt).to(DEV) for k in ks]
    bias = torch.randn(512, generator=g).float()
    scale, shift = torch.rand(512, generator=g) * 20 + 5, torch.randn(512, generator=g) * 0.1
    res = torch.randn(m, 512, generator=g).to(dt).to(DEV) if residual else None
    out = Activation(m, 512, precision, DEV)
    segs = [(a.data_ptr(), k, b.data_ptr(), k, k) for a, b, k in zip(As, Bs, ks)]
    extra = {}
    if pool:                                   # pool-fused epilogue: block sums instead of rows (1 block in 125 is kept)
        nb = (m + 31) // 32
        sums = torch.empty(nb, 512, dtype=torch.float32, device=DEV)
        keep = (torch.arange(nb) % 125 == 0).to(torch.uint8).to(DEV)
        extra = dict(pool_block_sums=sums.data_ptr(), pool_block_keep=keep.data_ptr())
    def run():
        engine.gemm512(segs, m, precision, out, bias=bias.data_ptr(), bn_scale=scale.data_ptr() if normalize else None,
                       bn_shift=shift.data_ptr() if normalize else None, residual=engine._p(res), ldr=512,
                       normalize=normalize, relu=normalize, **extra)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * m * sum(ks) * 512
    line = f"{name:34s} m={m} K={sum(ks)} {precision}: {ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s"
    lib = capi.load()
    if hasattr(lib, "bg_gemm_prof_host"):
        n = 148 * 8
        buf = (C.c_ulonglong * n)()
        lib.bg_gemm_prof_host(buf, n)
        v = torch.tensor(list(buf), dtype=torch.float64).view(148, 8)
        lead = v[0::2]                       # leader CTAs own the MMA counters
        f = lambda t: f"{t.mean().item() / 1e3:.0f}k"
        line += (f"\n      cycles/CTA: producer wait-empty {f(v[:, 0])} | mma wait-full {f(lead[:, 1])} "
                 f"wait-tmem {f(lead[:, 2])} total {f(lead[:, 3])} | epi wait-full {f(v[:, 4])} pass1 {f(v[:, 5])} "
                 f"pass2 {f(v[:, 6])} total {f(v[:, 7])}")
    print(line, flush=True)


if __name__ == "__main__":
    capi.device_check()
    M = 1030398
    case("sage update, full epilogue + skip", M, [512, 512], "fp16", True, True)
    case("sage update, full epilogue", M, [512, 512], "fp16", True, False)
    case("sage update, pool-fused epilogue", M, [512, 512], "fp16", True, False, pool=True)
    case("sage update, bias only", M, [512, 512], "fp16", False, False)
    case("encoder 128->512, bias only", M, [128], "fp16", False, False)
    case("sage update bf16 + skip", M, [512, 512], "bf16", True, True)
    case("sage update tf32 + skip", M, [512, 512], "tf32", True, True, iters=5)
