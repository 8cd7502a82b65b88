#!/bin/bash
mkdir -p gpurun_out
python tools/agg_stiff_bench.py > gpurun_out/agg_stiff.log 2>&1; tail -2 gpurun_out/agg_stiff.log
ncu --set full --clock-control none --import-source on -k regex:'^k_aggregate_rows$' -s 3 -c 1 -o gpurun_out/prof_agg_stiff -f python tools/agg_stiff_bench.py > gpurun_out/ncu_agg_stiff.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_agg_stiff.ncu-rep
