#!/bin/bash
mkdir -p gpurun_out
NGB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 --csv --log-file gpurun_out/exp10_sweep_launches.csv python bench.py --workload sweep --steps 1 --warmup 0 > gpurun_out/exp10_ncu.log 2>&1
