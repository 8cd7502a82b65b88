#!/bin/bash
# Runs the given pytest selection on the GPU box, log under gpurun_out/.  usage: bash tools/gpu_one.sh <name> <pytest args...>
mkdir -p gpurun_out
name=$1; shift
timeout 900 python -m pytest "$@" -q -m gpu > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log
tail -n 60 gpurun_out/$name.log
