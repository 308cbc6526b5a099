#!/bin/bash
# gpu_variants.sh TAG [variant ...]: GPU parity tests on the in-tree build, then the Newton-step timing
# (tests/gpu_profile_run.py, 4096 samples) of the in-tree build and of every build/variants/<variant>/libngb200.so,
# then the launch list of the in-tree build.  NCU_FULL=1 adds a full capture of ngb_k_bsim4_load with source counters.
TAG=$1; shift
mkdir -p gpurun_out; L=gpurun_out/$TAG.log; : > $L
if [ -z "$SKIP_TESTS" ]; then ( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) >> $L; fi
cd tests
run() { echo "== $1" >> ../$L; shift; for i in 1 2; do ( env "$@" timeout 120 python gpu_profile_run.py 4096 ) 2>&1 | tail -1 >> ../$L; done; }
run "in-tree" A=1
for v in "$@"; do run "$v" NGB200_LIB=$PWD/../build/variants/$v/libngb200.so; done
cd ..
NGB_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python tests/gpu_profile_run.py 4096 > gpurun_out/${TAG}_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv >> $L 2>&1
if [ -n "$NCU_FULL" ]; then
  NGB_NO_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:ngb_k_bsim4_load -s 30 -c 1 -f -o gpurun_out/${TAG}_b4load python tests/gpu_profile_run.py 4096 > gpurun_out/${TAG}_b4load.log 2>&1
  ncu -i gpurun_out/${TAG}_b4load.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_b4load_source.csv 2>/dev/null
fi
cat $L
