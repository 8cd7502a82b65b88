#!/bin/bash
# Runs each GPU test group in its own process (a trapped kernel poisons the CUDA context
# of its process only) and collects logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
run probe_cg2 python tools/gemm_probe.py 2
run t_csr python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "csr or graph_ptr"
run t_enc python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "encoder"
run t_agg python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "aggregate"
run t_gemm python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm"
run t_pool python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "pool"
run t_fwd python -m pytest tests/test_gpu_forward.py -q -m gpu
run smoke python __graft_entry__.py --smoke
run bench python bench.py --steps 10 --warmup 3
