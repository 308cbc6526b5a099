#!/bin/bash
mkdir -p gpurun_out
cap() { # name kernel env...
  n=$1; k=$2; shift 2
  env "$@" NGB_NO_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 30 -c 1 -f -o gpurun_out/$n python tests/gpu_profile_run.py 4096 > gpurun_out/$n.log 2>&1
}
cap x3_mono ngb_k_bsim4_load A=1
cap x3_core ngb_k_b4_core NGB_B4_SPLIT=1
cap x3_fin ngb_k_b4_fin NGB_B4_SPLIT=1
cap x3_lu ngb_k_lu_packed A=1
ls -la gpurun_out/*.ncu-rep
