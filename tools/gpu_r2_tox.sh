#!/bin/bash
# GPU parity tests, then the default bench (continuous per-sample toxe, field-major rows) and the same with row-major rows
mkdir -p gpurun_out; L=gpurun_out/r2_tox.log; : > $L
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) >> $L
( time timeout 1200 python bench.py --steps 1 ) > gpurun_out/r2_tox_bench.json 2>> $L
tail -c 2500 gpurun_out/r2_tox_bench.json >> $L
( time NGB_BENCH_ROWMAJOR=1 timeout 1200 python bench.py --steps 1 --warmup 3 ) > gpurun_out/r2_tox_rowmajor_bench.json 2>> $L
tail -c 1200 gpurun_out/r2_tox_rowmajor_bench.json >> $L
( time timeout 900 python bench.py --workload sweep ) > gpurun_out/r02_bench_sweep.json 2>> $L
tail -c 1500 gpurun_out/r02_bench_sweep.json >> $L
cat $L
