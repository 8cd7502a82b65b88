#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=30 run t_round2 python -m pytest tests/test_gpu_round2.py tests/test_gpu_forward.py -q -m gpu
