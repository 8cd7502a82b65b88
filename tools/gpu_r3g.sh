#!/bin/bash
mkdir -p gpurun_out
python tools/agg128_bench.py > gpurun_out/agg128.log 2>&1; cat gpurun_out/agg128.log
ncu --set full --clock-control none --import-source on -k regex:"k_aggregate_rows128" -s 2 -c 1 -o gpurun_out/prof_agg128 -f python tools/agg128_bench.py 3 > gpurun_out/ncu_agg128.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_agg128.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"k_aggregate_rows<" -s 2 -c 1 -o gpurun_out/prof_agg512 -f python tools/agg_bench.py > gpurun_out/ncu_agg512.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_agg512.ncu-rep
