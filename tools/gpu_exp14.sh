#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/exp14.log; : > $L
( timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -2 >> $L
cd tests
run() { echo "== $1" >> ../$L; shift; ( env "$@" timeout 120 python gpu_profile_run.py 4096 ) 2>&1 | tail -1 >> ../$L; }
run "stamps at computed rows" A=1
run "stamps through the row table" NGB200_LIB=$PWD/../build/variants/lookup/libngb200.so
run "computed rows again" A=1
cd ..
NGB_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/exp14_launches.csv python tests/gpu_profile_run.py 4096 > gpurun_out/exp14_ncu.log 2>&1
cat $L
