#!/bin/bash
# aggregate: stream-ring geometry vs L1 carve-out (r02 session 3)
mkdir -p gpurun_out
echo "=== product"; timeout 300 python tools/agg_bench.py 2>&1 | tail -4
for so in buckgnn_b200/lib/variants/a_*.so; do
  n=$(basename $so .so)
  echo "=== $n"
  BG_LIB_PATH=$PWD/$so timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "aggreg" 2>&1 | tail -1
  BG_LIB_PATH=$PWD/$so timeout 300 python tools/agg_bench.py 2>&1 | tail -4
done
