#!/bin/bash
# ncu evidence for the bench command (run under gpurun, 1 GPU). Outputs under gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_gemm512|k_aggregate_rows|k_aggregate_hubs|k_hub_finalize|k_encoder_front|k_pool_partial" -s 40 -c 20 -o gpurun_out/prof_top -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu_full.log
