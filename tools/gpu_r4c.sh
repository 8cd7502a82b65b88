#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
run t_all python -m pytest tests -q -x -m gpu
TAILN=5 run agg python tools/agg_bench.py
TAILN=2 run bench python bench.py --steps 10 --warmup 3 --no-extras
TAILN=2 run bench_configs python tools/bench_configs.py cfg5
TAILN=3 run sweep1 python tools/bench_sweep.py --graphs 8000
