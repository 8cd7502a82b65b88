#!/bin/bash
# ncu evidence for the bench command (run under gpurun, 1 GPU). Outputs under gpurun_out/ -- summaries are made on the
# box and the raw reports are dropped when they would not fit gpurun's 64 MiB return limit.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_gemm512|k_aggregate_rows|k_aggregate_hubs|k_hub_finalize|k_encoder_front|k_pool_partial" -s 40 -c 20 -o gpurun_out/prof_top -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu_full.log
python tools/ncu_summary.py gpurun_out/prof_top.ncu-rep gpurun_out/ncu_full_summary.csv > gpurun_out/ncu_full_summary.txt 2>&1
if [ "$1" == "sag" ]; then
  SAG="python tools/bench_configs.py sag1"
  $SAG > gpurun_out/sag_plain.log 2>&1 &&
  ncu --set full --clock-control none -k regex:"k_sag_|k_gather_rows" -c 11 -o gpurun_out/prof_sag -f $SAG > gpurun_out/ncu_sag_full.log 2>&1
  echo "sag full capture exit $?"
  python tools/ncu_summary.py gpurun_out/prof_sag.ncu-rep gpurun_out/ncu_sag_full_summary.csv > gpurun_out/ncu_sag_full_summary.txt 2>&1
  rm -f gpurun_out/prof_sag.ncu-rep
fi
du -sh gpurun_out
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/prof_top.ncu-rep; echo "dropped prof_top.ncu-rep (too large to return)"; fi
