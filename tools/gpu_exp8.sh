#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/exp8.log; : > $L
cd tests
run() { echo "== $1" >> ../$L; shift; ( env "$@" timeout 120 python gpu_profile_run.py 4096 ) 2>&1 | tail -1 >> ../$L; }
run "default" A=1
for v in bar bar512 c512; do run "$v" NGB200_LIB=$PWD/../build/variants/$v/libngb200.so; done
run "default" A=1
cd ..; cat $L
