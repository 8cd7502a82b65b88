#!/bin/bash
# GPU visit for the SAGPooling variants: the new tests first (own process), then the whole GPU suite + smoke + bench,
# the sag bench and its ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; }
TAILN=40 run t_sag python -m pytest tests/test_gpu_sag.py -q -m gpu
run t_all python -m pytest tests -q -x -m gpu --deselect tests/test_gpu_sag.py
run smoke python __graft_entry__.py --smoke
TAILN=3 run bench python bench.py --steps 10 --warmup 3
TAILN=6 run bench_sag python tools/bench_configs.py sag
# launch list of the GraphSAGE_SAG forwards (the first case of the sag bench: 2 warm-up + 5 timed forwards)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/sag_launches.csv python tools/bench_configs.py sag1 > gpurun_out/ncu_sag.log 2>&1
echo "sag launch list exit $?"
