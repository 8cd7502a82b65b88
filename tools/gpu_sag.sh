#!/bin/bash
# GPU visit for the SAGPooling variants: the new tests first (own process), then the whole GPU suite + smoke + bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; }
TAILN=40 run t_sag python -m pytest tests/test_gpu_sag.py -q -m gpu
run t_all python -m pytest tests -q -x -m gpu --deselect tests/test_gpu_sag.py
run smoke python __graft_entry__.py --smoke
TAILN=3 run bench python bench.py --steps 10 --warmup 3
TAILN=5 run bench_sag python tools/bench_configs.py sag
