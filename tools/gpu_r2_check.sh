#!/bin/bash
# gpu_r2_check.sh TAG: GPU parity tests, Newton-step timing of the in-tree build (specialised and generic BSIM4 load),
# in-situ stage breakdown, launch list
TAG=$1; shift
mkdir -p gpurun_out; L=gpurun_out/$TAG.log; : > $L
if [ -z "$SKIP_TESTS" ]; then ( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) >> $L; fi
cd tests
echo "== specialised" >> ../$L
for i in 1 2; do ( timeout 120 python gpu_profile_run.py 4096 ) 2>&1 | tail -1 >> ../$L; done
( timeout 120 python gpu_profile_run.py 4096 stages ) 2>&1 | tail -1 >> ../$L
echo "== generic" >> ../$L
( NGB_B4_GENERIC=1 timeout 120 python gpu_profile_run.py 4096 stages ) 2>&1 | tail -2 >> ../$L
for v in "$@"; do echo "== $v" >> ../$L; ( NGB200_LIB=$PWD/../build/variants/$v/libngb200.so timeout 120 python gpu_profile_run.py 4096 stages ) 2>&1 | tail -2 >> ../$L; done
cd ..
if [ -z "$SKIP_NCU" ]; then
NGB_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python tests/gpu_profile_run.py 4096 > gpurun_out/${TAG}_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv >> $L 2>&1
fi
cat $L
