#!/bin/bash
# gpu_r2_arrtran.sh: config 4 with the LU at 48x48, 64x64, 128x128 (DC op + transient, grid-wide LU)
mkdir -p gpurun_out; L=gpurun_out/r2_arrtran.log; : > $L
for c in 2304 4096 16384; do
  echo "== cells $c" >> $L
  ( timeout 900 python bench.py --workload array_tran --cells $c --steps 1 --warmup 3 2>&1 | tail -1 ) > gpurun_out/r02_bench_array_tran_$c.json
  head -c 3000 gpurun_out/r02_bench_array_tran_$c.json >> $L; echo >> $L
done
cat $L
