#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/exp2.log; : > $L
( time NGB_B4_SPLIT=1 timeout 900 python -m pytest tests -m gpu -x -q ) >> $L 2>&1
cd tests
run() { echo "== $1" >> ../$L; shift; ( env "$@" timeout 120 python gpu_profile_run.py 4096 ) 2>&1 | tail -3 >> ../$L; }
run "default (branches on, one-kernel load)" A=1
run "branches off" NGB_NO_BRANCH=1
run "split load, 128 regs" NGB_B4_SPLIT=1
run "split load again" NGB_B4_SPLIT=1
run "split p2c3" NGB_B4_SPLIT=1 NGB200_LIB=$PWD/../build/variants/p2c3/libngb200.so
run "split p3" NGB_B4_SPLIT=1 NGB200_LIB=$PWD/../build/variants/p3/libngb200.so
cd ..
NGB_B4_SPLIT=1 NGB_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/exp2_launches.csv python tests/gpu_profile_run.py 4096 > gpurun_out/exp2_ncu.log 2>&1
tail -40 $L
