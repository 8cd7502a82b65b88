#!/bin/bash
# One GPU-box visit: the driver's checks (pytest -m gpu, smoke, bench) + the training bench + the ncu evidence.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
run t_all python -m pytest tests -q -x -m gpu
run smoke python __graft_entry__.py --smoke
run bench python bench.py --steps 10 --warmup 3
run bench_train python tools/bench_train.py --steps 10 --warmup 3
run bench_train_bf16 python tools/bench_train.py --steps 10 --warmup 3 --precision bf16
TAILN=3 run bench_configs python tools/bench_configs.py
if [ "$1" != "noprof" ]; then
  bash tools/profile.sh
  # launch list of training steps (cfg 4): two steps after three warm-up steps, skipping the warm-up launches
  TRAIN="python tools/bench_train.py --steps 2 --warmup 3"
  $TRAIN > gpurun_out/train_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 700 --csv --log-file gpurun_out/train_launches.csv $TRAIN > gpurun_out/ncu_train.log 2>&1
  echo "train launch list exit $?"
fi
