#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=3 run t_train python -m pytest tests/test_gpu_train.py tests/test_gpu_round2.py tests/test_gpu_poolfuse.py -q -x -m gpu
TAILN=1 run bench_train python tools/bench_train.py --steps 20 --warmup 5
TAILN=1 run bench_train_bf16 python tools/bench_train.py --steps 20 --warmup 5 --precision bf16
