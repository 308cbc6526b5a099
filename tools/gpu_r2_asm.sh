#!/bin/bash
# gpu_r2_asm.sh: tiled assembly (matrix values through shared memory, coalesced Ax stores) against the direct one;
# BSIM4trunc with a CTA per sample against a warp per sample
mkdir -p gpurun_out; L=gpurun_out/r2_asm.log; : > $L
( NGB_ASM_TILED=1 NGB_LTE_CTA=1 timeout 900 python -m pytest tests/test_load_parity.py tests/test_tran_parity.py tests/test_lu_parity.py tests/test_synth_array.py tests/test_bench_parity.py -m gpu -x -q 2>&1 | tail -3 ) >> $L
cd tests
for v in "1 1" "0 1" "1 0" "0 0" "1 1"; do set -- $v; echo "== NGB_ASM_TILED=$1 NGB_LTE_CTA=$2" >> ../$L; ( NGB_ASM_TILED=$1 NGB_LTE_CTA=$2 timeout 120 python gpu_profile_run.py 4096 ) 2>&1 | tail -1 >> ../$L; done
for v in "1 1" "0 0"; do set -- $v; echo "== NGB_ASM_TILED=$1 NGB_LTE_CTA=$2 stages" >> ../$L; ( NGB_ASM_TILED=$1 NGB_LTE_CTA=$2 timeout 120 python gpu_profile_run.py 4096 stages ) 2>&1 | tail -1 >> ../$L; done
cd ..
cat $L
