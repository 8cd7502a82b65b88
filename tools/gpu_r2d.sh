#!/bin/bash
mkdir -p gpurun_out
python tools/bench_train.py --steps 20 --warmup 5 > gpurun_out/bench_train.log 2>&1; tail -1 gpurun_out/bench_train.log
python -c "
import cProfile, pstats, sys, io
sys.argv=['bench_train.py','--steps','30','--warmup','5']
sys.path.insert(0,'tools')
import runpy
pr=cProfile.Profile(); pr.enable()
runpy.run_path('tools/bench_train.py', run_name='__main__')
pr.disable()
s=io.StringIO(); pstats.Stats(pr,stream=s).sort_stats('tottime').print_stats(45); print(s.getvalue())
" > gpurun_out/train_cprofile.log 2>&1
head -80 gpurun_out/train_cprofile.log
