#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/exp4.log; : > $L
cd tests
run() { echo "== $1" >> ../$L; shift; ( env "$@" timeout 120 python gpu_profile_run.py 4096 ) 2>&1 | tail -1 >> ../$L; }
run "one-kernel load" A=1
run "split" NGB_B4_SPLIT=1
for v in i_all i_core i_small i_corefin f3; do run "split $v" NGB_B4_SPLIT=1 NGB200_LIB=$PWD/../build/variants/$v/libngb200.so; done
run "one-kernel i_all" NGB200_LIB=$PWD/../build/variants/i_all/libngb200.so
cd ..
NGB200_LIB=$PWD/build/variants/i_all/libngb200.so NGB_B4_SPLIT=1 NGB_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/exp4_launches.csv python tests/gpu_profile_run.py 4096 > gpurun_out/exp4_ncu.log 2>&1
cat $L
