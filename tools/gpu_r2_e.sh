#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/r2_e.log; : > $L
( timeout 600 python -m pytest tests/test_tran_parity.py -m gpu -q -k "tox" 2>&1 | tail -3 ) >> $L
for mode in "NGB_BENCH_FIELDMAJOR=1" "A=1"; do
  echo "== bench $mode" >> $L
  ( env $mode timeout 900 python bench.py --steps 1 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'load_ms', d['roofline']['avg_launch_ms'], 'parity', d['parity_check']['ok'], d['config']['layout'])" ) >> $L 2>&1
done
cat $L
