#!/bin/bash
# A/B of GEMM variants: tools/gemm_bench.py (and the GEMM parity tests) under each library in buckgnn_b200/lib/variants
mkdir -p gpurun_out
for so in buckgnn_b200/lib/variants/*.so; do
  n=$(basename $so .so)
  echo "=== $n"
  BG_LIB_PATH=$PWD/$so timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py -q -m gpu -k "gemm or eagnn or mean_6x512 or other_graphsage" > gpurun_out/t_$n.log 2>&1; echo "tests exit $?"; tail -1 gpurun_out/t_$n.log
  BG_LIB_PATH=$PWD/$so timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_$n.log 2>&1; echo "exit $?"
  cat gpurun_out/gemm_$n.log | head -${LINES_PER:-8}
done
