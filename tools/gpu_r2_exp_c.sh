#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/r2_exp_c.log; : > $L
for cells in 20000 500000; do
  echo "== array cells=$cells" >> $L
  ( timeout 600 python bench.py --workload array --cells $cells --impl ours 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'load_ms', d['roofline']['avg_launch_ms'], 'evals', d['roofline']['units_per_launch'], 'ns_per_eval', 1e6*d['roofline']['avg_launch_ms']/d['roofline']['units_per_launch'], 'frac', d['roofline']['frac'])" ) >> $L 2>&1
done
cd tests
for cta in 512 1024; do echo "== ro101 NGB_LU_CTA=$cta" >> ../$L; ( NGB_LU_CTA=$cta timeout 300 python gpu_profile_ro101.py ) 2>&1 | tail -2 >> ../$L; done
cd ..
for mode in "NGB_BENCH_TOX_SIGMA=1e-9" "NGB_BENCH_TOX_SIGMA=1e-9 NGB_BENCH_ROWMAJOR=1"; do
  echo "== bench $mode" >> $L
  ( env $mode timeout 900 python bench.py --steps 1 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'load_ms', d['roofline']['avg_launch_ms'], 'parity', d['parity_check']['ok'], d['config']['layout'])" ) >> $L 2>&1
done
cat $L
