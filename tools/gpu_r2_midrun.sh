#!/bin/bash
# gpu_r2_midrun.sh [variant ...]: (1) Newton-step timing + in-situ stage breakdown of the in-tree build and of every
# build/variants/<variant>/libngb200.so (tests/gpu_profile_run.py, 4096 samples); (2) one `ncu --set full` capture of
# ngb_k_bsim4_load at Newton step 15 000 of the default bench's transient, i.e. in the regime the bench averages over
# (every sample at its own phase, time step and state-ring position), with the source-level counters exported
mkdir -p gpurun_out; L=gpurun_out/r2_mid.log; : > $L
cd tests
run() { echo "== $1" >> ../$L; shift; ( env "$@" timeout 120 python gpu_profile_run.py 4096 ) 2>&1 | tail -1 >> ../$L
        ( env "$@" timeout 120 python gpu_profile_run.py 4096 stages ) 2>&1 | tail -2 >> ../$L; }
run "in-tree" A=1
for v in "$@"; do run "$v" NGB200_LIB=$PWD/../build/variants/$v/libngb200.so; done
run "in-tree again" A=1
cd ..
if [ -z "$SKIP_NCU" ]; then
NGB_NO_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:ngb_k_bsim4_load -s 15000 -c 1 --kill 1 -f \
  -o gpurun_out/r02_b4load_mid python bench.py --steps 1 --warmup 1 > gpurun_out/r2_mid_ncu.log 2>&1
ncu -i gpurun_out/r02_b4load_mid.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_b4load_mid_source.csv 2>/dev/null
ncu -i gpurun_out/r02_b4load_mid.ncu-rep --page details > gpurun_out/r02_b4load_mid_details.txt 2>/dev/null
ncu -i gpurun_out/r02_b4load_mid.ncu-rep --page raw --csv > gpurun_out/r02_b4load_mid_raw.csv 2>/dev/null
tail -3 gpurun_out/r2_mid_ncu.log >> $L
fi
cat $L
