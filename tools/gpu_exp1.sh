#!/bin/bash
# one GPU call: parity tests, then the Newton step of the 4096-sample batch under each experiment build
mkdir -p gpurun_out; L=gpurun_out/exp1.log; : > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $L
( time timeout 900 python -m pytest tests -m gpu -x -q ) >> $L 2>&1
cd tests
run() { echo "== $1" >> ../$L; shift; ( env "$@" timeout 120 python gpu_profile_run.py 4096 ) >> ../$L 2>&1; }
run "default (pk2 LU)" A=1
run "default again" A=1
run "LU first packing" NGB_LU_V1=1
for v in r80 r96 r168 r128c128; do run "variant $v" NGB200_LIB=$PWD/../build/variants/$v/libngb200.so; done
cd ..
NGB_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/exp1_launches.csv python tests/gpu_profile_run.py 4096 > gpurun_out/exp1_ncu.log 2>&1
tail -50 $L
