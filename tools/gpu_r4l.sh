#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_poolfuse.py tests/test_golden.py tests/test_gpu_fullsize.py -q -x -m gpu 2>&1 | tail -2
python tools/layer_times_probe.py 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit $?"
