#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/exp6.log; : > $L
( time timeout 900 python -m pytest tests -m gpu -x -q ) >> $L 2>&1
cd tests
run() { echo "== $1" >> ../$L; shift; ( env "$@" timeout 120 python gpu_profile_run.py 4096 ) 2>&1 | tail -1 >> ../$L; }
run "tiled assembly" A=1
run "thread-per-target assembly" NGB_ASM_TILED=0
run "tiled again" A=1
cd ..
echo "== sweep, branches on" >> $L; ( timeout 600 python bench.py --workload sweep --steps 2 --warmup 1 ) 2>&1 | tail -1 >> $L
echo "== sweep, branches off" >> $L; ( NGB_NO_BRANCH=1 timeout 600 python bench.py --workload sweep --steps 2 --warmup 1 ) 2>&1 | tail -1 >> $L
NGB_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/exp6_launches.csv python tests/gpu_profile_run.py 4096 > gpurun_out/exp6_ncu.log 2>&1
cat $L
