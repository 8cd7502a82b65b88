#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=15 run t_pool python -m pytest tests/test_gpu_poolfuse.py -q -x -m gpu
run t_all python -m pytest tests -q -x -m gpu
TAILN=2 run bench python bench.py --steps 10 --warmup 3 --no-extras
