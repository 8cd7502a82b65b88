#!/bin/bash
# Multi-GPU evidence, round 2 final code: gpurun --gpus N -- bash tools/gpu_multi3.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_${N}.log 2>&1; echo "bench exit $?"; grep '^{' gpurun_out/scale_${N}.log | cut -c1-300
timeout 600 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/scale_ref_${N}.log 2>&1; echo "bench ref exit $?"; grep '^{' gpurun_out/scale_ref_${N}.log | cut -c1-200
timeout 600 $TR tools/bench_sweep.py --gpus $N --graphs ${SWEEP_GRAPHS:-16000} > gpurun_out/sweep_${N}gpu.log 2>&1; echo "sweep exit $?"; grep '^{' gpurun_out/sweep_${N}gpu.log
timeout 600 $TR tools/train_ddp_check.py > gpurun_out/ddp_check_${N}gpu.log 2>&1; echo "ddp check exit $?"; tail -3 gpurun_out/ddp_check_${N}gpu.log
