#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/exp12.log; : > $L
( timeout 900 python -m pytest tests/test_lu_parity.py tests/test_variants.py tests/test_tran_parity.py -m gpu -x -q ) 2>&1 | tail -2 >> $L
cd tests
run() { echo "== $1" >> ../$L; shift; ( env "$@" timeout 120 python gpu_profile_run.py 4096 ) 2>&1 | tail -1 >> ../$L; }
run "early runs" A=1
run "no hoisting" NGB_LU_NOHOIST=1
run "early runs again" A=1
cd ..
NGB_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/exp12_launches.csv python tests/gpu_profile_run.py 4096 > gpurun_out/exp12_ncu.log 2>&1
cat $L
