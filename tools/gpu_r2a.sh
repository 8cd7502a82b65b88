#!/bin/bash
# round 2, first GPU visit: full GPU test suite with the round-2 tests, the EA-GNN precision probe, a baseline bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=40 run t_round2 python -m pytest tests/test_gpu_round2.py tests/test_golden.py -q -m gpu
TAILN=20 run probe_eagnn python tools/eagnn_precision_probe.py
TMO=1500 TAILN=15 run t_all python -m pytest tests -q -m gpu --deselect tests/test_gpu_round2.py --deselect tests/test_golden.py
run smoke python __graft_entry__.py --smoke
TAILN=2 run bench python bench.py --steps 10 --warmup 3
