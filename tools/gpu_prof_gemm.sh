#!/bin/bash
# ncu --set full with source correlation on the SAGE update GEMM (one warm launch), plus the role-cycle profile
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_gemm512" -s 6 -c 2 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture exit $?"; tail -2 gpurun_out/ncu_gemm.log
ls -la gpurun_out/*.ncu-rep
