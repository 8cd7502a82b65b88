#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=40 run corun python tools/corun_probe.py
TAILN=1 run bench python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1]
d=json.loads(l); print(d['ms_per_step'], d['kernel_ms_per_step'], d['e2e']['value'])
PY
