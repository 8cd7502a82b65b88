#!/bin/bash
# GEMM probes (r02 session 3): operand-traffic and idle-wait probes, each under tools/gemm_bench.py with the SM clock
# and board power sampled every 100 ms beside it
mkdir -p gpurun_out
for so in buckgnn_b200/lib/variants/*.so; do
  n=$(basename $so .so)
  echo "=== $n"
  nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.active --format=csv,noheader -lms 100 > gpurun_out/smi_$n.csv 2>&1 &
  SMI=$!
  BG_LIB_PATH=$PWD/$so timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_$n.log 2>&1; echo "exit $?"
  kill $SMI
  head -${LINES_PER:-4} gpurun_out/gemm_$n.log
  sort gpurun_out/smi_$n.csv | uniq -c | sort -rn | head -5
done
