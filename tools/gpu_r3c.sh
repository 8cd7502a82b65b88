#!/bin/bash
mkdir -p gpurun_out
for so in buckgnn_b200/lib/variants/*.so; do
  n=$(basename $so .so)
  echo "=== $n"
  BG_LIB_PATH=$PWD/$so timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_$n.log 2>&1; echo "exit $?"
  cat gpurun_out/gemm_$n.log | head -${LINES_PER:-8}
done
