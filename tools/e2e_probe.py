"""Where does the end-to-end step time go?  Runs the pipelined inference loop of bench.py with per-step event
timestamps and prints: period, forward span, gap between forwards, H2D copy span and its start offset."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buckgnn_b200.model import BuckGNN
from buckgnn_b200.pipeline import PipelinedInference
from buckgnn_b200.synth import config_batch

DEV = "cuda:0"


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    torch.manual_seed(0)
    model = BuckGNN(16, 5, 512, 6, "mean", model_name="GraphSage_meanAggr").to(DEV).eval()
    host = config_batch(1).pin_memory()
    for depth in (1, 2):
        for with_copy in (True, False):
            evs = {}
            host_t = {}

            def hook(step, before):
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                evs.setdefault(step, []).append(ev)
                host_t.setdefault(step, []).append(time.perf_counter())
            pipe = PipelinedInference(model, (host for _ in range(steps)), DEV, depth=depth, on_launch=hook)
            pipe.prefetcher.time_copies = True
            if not with_copy:                      # same loop, but the "copies" move nothing (resident staging)
                res = host.to(DEV)
                pipe.prefetcher._stage = lambda h, k: res
            t0 = time.perf_counter()
            n = sum(1 for _ in pipe)
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3 / n
            spans = [evs[i][0].elapsed_time(evs[i][1]) for i in range(n)]
            gaps = [evs[i][1].elapsed_time(evs[i + 1][0]) for i in range(n - 1)]
            hostcall = [(host_t[i][1] - host_t[i][0]) * 1e3 for i in range(n)]
            med = lambda v: sorted(v)[len(v) // 2]
            line = dict(depth=depth, h2d=with_copy, wall_ms_per_step=round(wall, 3), fwd_span_med=round(med(spans), 3),
                        fwd_span_mean=round(sum(spans[2:]) / len(spans[2:]), 3), gap_med=round(med(gaps), 3),
                        gap_mean=round(sum(gaps[2:]) / len(gaps[2:]), 3), host_in_model_call_med=round(med(hostcall), 3))
            if with_copy:
                ce = pipe.prefetcher.copy_events
                line["copy_med"] = round(med([a.elapsed_time(b) for a, b in ce]), 3)
            print(line, flush=True)


if __name__ == "__main__":
    main()
