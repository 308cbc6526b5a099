#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/r2_exp_b.log; : > $L
cd tests
for cta in 256 128 64 32; do echo "== ro101 NGB_LU_CTA=$cta" >> ../$L; ( NGB_LU_CTA=$cta timeout 300 python gpu_profile_ro101.py ) 2>&1 | tail -2 >> ../$L; done
cd ..
for cells in 20000 50000 125000 250000 500000; do
  echo "== array cells=$cells" >> $L
  ( timeout 600 python bench.py --workload array --cells $cells --impl ours 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'load_ms', d['roofline']['avg_launch_ms'], 'evals', d['roofline']['units_per_launch'], 'ns_per_eval', 1e6*d['roofline']['avg_launch_ms']/d['roofline']['units_per_launch'])" ) >> $L 2>&1
done
cd tests
for S in 4096 16384 32768; do echo "== mc S=$S" >> ../$L; ( timeout 300 python gpu_profile_run.py $S stages ) 2>&1 | tail -2 >> ../$L; done
cd ..
cat $L
