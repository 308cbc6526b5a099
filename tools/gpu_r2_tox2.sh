#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/r2_tox2.log; : > $L
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) >> $L
( time timeout 1200 python bench.py --steps 1 ) > gpurun_out/r2_tox2_bench.json 2>> $L
tail -c 2200 gpurun_out/r2_tox2_bench.json >> $L
( time timeout 600 python bench.py --workload array ) > gpurun_out/r02_bench_array.json 2>> $L
tail -c 1300 gpurun_out/r02_bench_array.json >> $L
( time NGB_ARRAY_RANDOM_BIAS=1 timeout 600 python bench.py --workload array ) > gpurun_out/r02_bench_array_random.json 2>> $L
tail -c 700 gpurun_out/r02_bench_array_random.json >> $L
cat $L
