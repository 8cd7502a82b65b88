"""GPU debugging aid: run bg_gemm512 on structured inputs and print where it deviates."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buckgnn_b200 import capi, engine
from buckgnn_b200.engine import Activation

DEV = "cuda:0"


def run(m, k, precision, cg, pattern):
    a_dt, b_dt = (engine._TORCH[c] for c in engine.PRECISION_FORMATS[precision])
    g = torch.Generator().manual_seed(0)
    if pattern == "ones":
        a = torch.ones(m, k); b = torch.ones(512, k)
    elif pattern == "rowid":          # out[m, n] = m (A row = m/k everywhere, B ones)
        a = (torch.arange(m).float()[:, None] % 64).expand(m, k) / 64; b = torch.ones(512, k)
    elif pattern == "colid":          # out[m, n] = n % 64
        a = torch.ones(m, k); b = (torch.arange(512).float()[:, None] % 64).expand(512, k) / 64
    elif pattern == "kdelta":         # A = e_j rows -> out[m, n] = B[n, m % k]
        a = torch.zeros(m, k); a[torch.arange(m), torch.arange(m) % k] = 1; b = torch.randn(512, k, generator=g)
    else:
        a = torch.randn(m, k, generator=g); b = torch.randn(512, k, generator=g)
    a, b = a.to(a_dt), b.to(b_dt)
    want = (a.double() @ b.double().T).float()
    ad, bd = a.to(DEV).contiguous(), b.to(DEV).contiguous()
    out = Activation(m, 512, precision, DEV)
    out.data.fill_(-777.0)
    engine.gemm512([(ad.data_ptr(), k, bd.data_ptr(), k, k)], m, precision, out, cta_group=cg)
    torch.cuda.synchronize()
    got = out.data.float().cpu()
    err = (got - want).abs()
    tol = 0.05 * want.abs().max().item() + 1e-2
    bad = err > tol
    print(f"[m={m} k={k} {precision} cg={cg} {pattern}] max_err={err.max().item():.4g} "
          f"bad={bad.sum().item()}/{bad.numel()} untouched={(got == -777).sum().item()}")
    if bad.any():
        rows = torch.nonzero(bad.any(1)).flatten()
        cols = torch.nonzero(bad.any(0)).flatten()
        print("   bad rows:", rows[:16].tolist(), "... count", rows.numel(), "| bad cols:", cols[:16].tolist(), "... count", cols.numel())
        r = rows[0].item()
        print("   got ", got[r, :8].tolist(), "\n   want", want[r, :8].tolist())
        print("   got[.,256:264]", got[r, 256:264].tolist(), "\n   want", want[r, 256:264].tolist())
    return not bad.any()


if __name__ == "__main__":
    capi.device_check()
    cgs = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "2").split(",")]
    ok = True
    for cg in cgs:
        for precision in ("bf16", "fp16", "tf32"):
            k = 64 if precision == "tf32" else 128
            for pattern in ("ones", "rowid", "colid", "kdelta", "rand"):
                ok &= run(256, k, precision, cg, pattern)
            ok &= run(1000, 512, precision, cg, "rand")
            ok &= run(148 * 256 * 2 + 77, 256, precision, cg, "rand")
    print("PROBE", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
