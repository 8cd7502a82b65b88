#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=3 run t_sel python -m pytest tests/test_gpu_forward.py tests/test_gpu_kernels.py -q -x -m gpu
TAILN=2 run bench python bench.py --steps 10 --warmup 3 --no-extras
TAILN=2 run bench2 python bench.py --steps 25 --warmup 5 --no-extras
TAILN=3 run sweep1 python tools/bench_sweep.py --graphs 8000
