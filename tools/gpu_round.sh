#!/bin/bash
# One GPU-box visit: the driver's checks (pytest -m gpu, smoke, bench) + the other benches + the ncu evidence.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
run t_all python -m pytest tests -q -x -m gpu
run smoke python __graft_entry__.py --smoke
TAILN=2 run bench python bench.py --steps 10 --warmup 3
TAILN=2 run bench25 python bench.py --steps 25 --warmup 5 --no-cpu-baseline
TAILN=2 run bench_ref python bench.py --impl reference --steps 3 --warmup 1
TAILN=2 run bench_tf32 python bench.py --steps 5 --warmup 3 --precision tf32 --no-cpu-baseline
TAILN=2 run bench_fp32 python bench.py --steps 5 --warmup 3 --precision fp32 --no-cpu-baseline
TAILN=2 run bench_train python tools/bench_train.py --steps 10 --warmup 3
TAILN=2 run bench_train_bf16 python tools/bench_train.py --steps 10 --warmup 3 --precision bf16
TAILN=4 run bench_configs python tools/bench_configs.py
TAILN=6 run bench_sag python tools/bench_configs.py sag
TAILN=2 run sweep1 python tools/bench_sweep.py --graphs 10000
TAILN=12 run gemm_bench python tools/gemm_bench.py
TAILN=5 run agg_bench python tools/agg_bench.py
if [ "$1" != "noprof" ]; then
  bash tools/profile.sh
  # launch list of training steps (cfg 4): two steps after three warm-up steps, skipping the warm-up launches
  TRAIN="python tools/bench_train.py --steps 2 --warmup 3"
  $TRAIN > gpurun_out/train_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 700 --csv --log-file gpurun_out/train_launches.csv $TRAIN > gpurun_out/ncu_train.log 2>&1
  echo "train launch list exit $?"
  # SAGPooling kernels: full capture of one forward's worth (fp16 storage case would be `sag` case 3; the first case is fp32)
  SAG="python tools/bench_configs.py sag1"
  $SAG > gpurun_out/sag_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"k_sag_|k_gather_rows" -c 11 -o gpurun_out/prof_sag -f $SAG > gpurun_out/ncu_sag_full.log 2>&1
  echo "sag full capture exit $?"
fi
