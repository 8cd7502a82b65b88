"""GPU micro-benchmark of bg_sage_aggregate alone on the cfg-2 batch (256 plates): ms per launch and
achieved algorithmic GB/s for the folded-hub and the generic hub path.  With --variants it builds the
library with different -D tuning macros and runs each in a subprocess (BG_LIB_PATH)."""
import os
import subprocess
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

VARIANTS = {
    "s2_l0": ["BG_AGG_STREAM_WARPS=2", "BG_AGG_STREAM_LEAD=0"],
    "s2_l16": ["BG_AGG_STREAM_WARPS=2", "BG_AGG_STREAM_LEAD=16"],
    "s2_l32": ["BG_AGG_STREAM_WARPS=2", "BG_AGG_STREAM_LEAD=32"],
    "s2_l48": ["BG_AGG_STREAM_WARPS=2", "BG_AGG_STREAM_LEAD=48"],
    "s2_l64": ["BG_AGG_STREAM_WARPS=2", "BG_AGG_STREAM_LEAD=64"],
    "s2_l96": ["BG_AGG_STREAM_WARPS=2", "BG_AGG_STREAM_LEAD=96"],
    "s1_l32": ["BG_AGG_STREAM_WARPS=1", "BG_AGG_STREAM_LEAD=32"],
    "s1_l64": ["BG_AGG_STREAM_WARPS=1", "BG_AGG_STREAM_LEAD=64"],
    "s3_l64": ["BG_AGG_STREAM_WARPS=3", "BG_AGG_STREAM_LEAD=64"],
    "s2_lm64": ["BG_AGG_STREAM_WARPS=2", "BG_AGG_STREAM_LEAD=-64"],
}


def build_variants():
    from buckgnn_b200 import build
    for name, defs in VARIANTS.items():
        out = os.path.join(build.LIB_DIR, f"libbuckgnn_b200_{name}.so")
        build.build(force=True, defines=defs, out=out)
        print("built", out)


def run_variants():
    from buckgnn_b200 import build
    for name in VARIANTS:
        out = os.path.join(build.LIB_DIR, f"libbuckgnn_b200_{name}.so")
        if os.path.exists(out):
            print("==", name, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=dict(os.environ, BG_LIB_PATH=out))


def main(iters=20):
    import torch
    from buckgnn_b200 import engine
    from buckgnn_b200.engine import Activation
    from buckgnn_b200.synth import config_batch
    dev = "cuda:0"
    b = config_batch(1)
    n = b.num_nodes
    idx = engine.build_graph_index(b.edge_index.to(dev), b.batch.to(dev), n)
    for precision in ("fp16", "tf32"):
        x, o = Activation(n, 512, precision, dev), Activation(n, 512, precision, dev)
        x.data.copy_(torch.randn(n, 512, device=dev))
        es = x.data.element_size()
        alg = 2 * n * 512 * es + 4 * idx.n_edges + 4 * (n + 1)
        for fold in (True, False):
            for _ in range(3):
                engine.aggregate(x, o, idx, "mean", fold_hubs=fold)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                engine.aggregate(x, o, idx, "mean", fold_hubs=fold)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            print(f"{precision} fold={fold}: {ms:.4f} ms  {alg / ms / 1e6:.0f} GB/s algorithmic", flush=True)


if __name__ == "__main__":
    if "--build" in sys.argv:
        build_variants()
    elif "--variants" in sys.argv:
        run_variants()
    else:
        main()
