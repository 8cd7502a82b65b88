#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=25 run t_agg python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "aggregate"
TAILN=6 run agg_window python tools/agg_bench.py
BG_AGG_WINDOW=0 TAILN=6 run agg_old python tools/agg_bench.py
TAILN=25 run t_fwd python -m pytest tests/test_gpu_forward.py tests/test_gpu_fullsize.py tests/test_gpu_round2.py tests/test_golden.py -q -m gpu -x
TAILN=1 run bench python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1]
d=json.loads(l); print(d['ms_per_step'], d['kernel_ms_per_step'], d['roofline_all']['aggregate']['frac'])
PY
TAILN=5 run cfg5 python tools/bench_configs.py cfg5
