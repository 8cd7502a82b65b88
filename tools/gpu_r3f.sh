#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=4 run t_sel python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py tests/test_gpu_fullsize.py tests/test_golden.py -q -m gpu -x
TAILN=1 run bench python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1]
d=json.loads(l); print(d['ms_per_step'], d['kernel_ms_per_step'], d['e2e']['value'], d['config']['parity_rel_err_vs_oracle_sample'], d['roofline']['frac'])
PY
