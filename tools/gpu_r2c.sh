#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=25 run t_new python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py tests/test_gpu_round2.py tests/test_gpu_train.py -q -m gpu -x
TAILN=3 run bench python bench.py --steps 10 --warmup 3
