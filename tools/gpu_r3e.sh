#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=2 run bench_full python bench.py --steps 10 --warmup 3
TAILN=4 run bench_configs python tools/bench_configs.py
TAILN=2 run bench_train_bf16 python tools/bench_train.py --steps 10 --warmup 3 --precision bf16
