#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/r2_d.log; : > $L
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) >> $L
( time timeout 600 python bench.py --workload array ) > gpurun_out/r02_bench_array.json 2>> $L
tail -c 1400 gpurun_out/r02_bench_array.json >> $L
( time timeout 1200 python bench.py ) > gpurun_out/r2_d_bench.json 2>> $L
tail -c 2300 gpurun_out/r2_d_bench.json >> $L
cat $L
