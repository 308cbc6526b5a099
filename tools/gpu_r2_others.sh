#!/bin/bash
# bench lines of the workloads the driver does not run, on the round's final build
mkdir -p gpurun_out; L=gpurun_out/r2_others.log; : > $L
for w in ro101 array sweep; do
  ( time timeout 600 python bench.py --workload $w ) > gpurun_out/r02_bench_$w.json 2>> $L
  tail -c 700 gpurun_out/r02_bench_$w.json >> $L; echo >> $L
done
cat $L
