#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'^k_aggregate_rows$' -s 2 -c 1 -o gpurun_out/prof_agg512 -f python tools/agg_bench.py > gpurun_out/ncu_agg512.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_agg512.ncu-rep; tail -3 gpurun_out/ncu_agg512.log
