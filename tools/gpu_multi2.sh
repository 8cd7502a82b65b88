#!/bin/bash
# N-GPU bench exactly as the driver launches it (torchrun, one rank per GPU)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.log 2>&1
echo "exit $?"; tail -n 3 gpurun_out/bench_${N}gpu.log
