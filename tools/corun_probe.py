"""GPU probe for the SM-partitioned SAGE layer: is the update GEMM power-bound (does it keep its
throughput on fewer SMs at a higher clock?), how does the aggregation scale with SMs, and what does a
layer cost when both run side by side on disjoint SMs (two streams, no hand-off yet)?

  python tools/corun_probe.py            -> one line per configuration
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buckgnn_b200 import capi, engine
from buckgnn_b200.engine import Activation
from buckgnn_b200.synth import config_batch

DEV = "cuda:0"


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    capi.device_check()
    lib = capi.load()
    b = config_batch(1)
    n = b.num_nodes
    idx = engine.build_graph_index(b.edge_index.to(DEV), b.batch.to(DEV), n)
    precision = "fp16"
    x, agg, out = (Activation(n, 512, precision, DEV) for _ in range(3))
    x.data.copy_(torch.relu(torch.randn(n, 512, device=DEV)) * 0.06)
    g = torch.Generator().manual_seed(0)
    wl = (torch.randn(512, 512, generator=g) / 22).half().to(DEV)
    wr = (torch.randn(512, 512, generator=g) / 22).half().to(DEV)
    bias = torch.randn(512, generator=g).float()
    scale, shift = torch.rand(512, generator=g) * 20 + 5, torch.randn(512, generator=g) * 0.1
    res = x

    def run_agg():
        engine.aggregate(x, agg, idx, "mean", fold_hubs=True)

    def run_gemm():
        segs = [(agg.data.data_ptr(), 512, wl.data_ptr(), 512, 512), (x.data.data_ptr(), 512, wr.data_ptr(), 512, 512)]
        engine.gemm512(segs, n, precision, out, bias=bias.data_ptr(), bn_scale=scale.data_ptr(), bn_shift=shift.data_ptr(),
                       residual=engine._p(res.data), ldr=512, normalize=True, relu=True)

    def part(gs, as_):
        capi._check(lib.bg_set_sm_partition(gs, as_), "bg_set_sm_partition")

    part(0, 0)
    t_a, t_g = timed(run_agg), timed(run_gemm)
    print(f"baseline: aggregate {t_a:.3f} ms, gemm {t_g:.3f} ms, serial layer {timed(lambda: (run_agg(), run_gemm())):.3f} ms", flush=True)
    for gs in (148, 128, 112, 104, 96, 88, 80, 64):
        part(gs, 0)
        print(f"gemm on {gs:3d} SMs: {timed(run_gemm):.3f} ms", flush=True)
    for as_ in (148, 96, 64, 52, 44, 36):
        part(0, as_)
        print(f"aggregate on {as_:3d} SMs: {timed(run_agg):.3f} ms", flush=True)
    # side by side on two streams (the data dependency is ignored here: this is the throughput probe)
    s_main = torch.cuda.current_stream()
    s_side = torch.cuda.Stream()
    for gs, as_ in ((104, 44), (96, 52), (88, 60), (80, 68), (112, 36)):
        part(gs, as_)

        def both():
            ev = torch.cuda.Event()
            ev.record(s_main)
            s_side.wait_event(ev)
            with torch.cuda.stream(s_side):
                run_agg()
            run_gemm()
            ev2 = torch.cuda.Event()
            ev2.record(s_side)
            s_main.wait_event(ev2)
        print(f"co-run gemm {gs} SMs || aggregate {as_} SMs: {timed(both):.3f} ms per layer", flush=True)
    part(0, 0)


if __name__ == "__main__":
    main()
