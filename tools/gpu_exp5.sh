#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/exp5.log; : > $L
cd tests
( timeout 120 python gpu_profile_run.py 4096 ) 2>&1 | tail -1 >> ../$L
for k in 1 2 3 4 8; do ( timeout 120 python gpu_concurrent_run.py 4096 $k ) 2>&1 | tail -2 >> ../$L; done
( NGB_NO_BRANCH=1 timeout 120 python gpu_concurrent_run.py 4096 2 ) 2>&1 | tail -2 >> ../$L
( timeout 120 python gpu_concurrent_run.py 8192 2 ) 2>&1 | tail -2 >> ../$L
( timeout 120 python gpu_concurrent_run.py 8192 4 ) 2>&1 | tail -2 >> ../$L
cd ..; cat $L
