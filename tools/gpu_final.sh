#!/bin/bash
# round-end measurement: GPU parity tests, smoke, the default bench line, the reference arm, other workloads, launch list, LU capture
mkdir -p gpurun_out; L=gpurun_out/final.log; : > $L
( time timeout 900 python -m pytest tests -m gpu -x -q ) >> $L 2>&1
python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "== bench default" >> $L; ( time timeout 1500 python bench.py ) >> $L 2>&1
echo "== bench reference arm" >> $L; ( time timeout 900 python bench.py --impl reference ) >> $L 2>&1
echo "== bench sweep" >> $L; ( timeout 900 python bench.py --workload sweep --steps 4 --warmup 3 ) 2>&1 | tail -1 >> $L
echo "== bench array" >> $L; ( timeout 900 python bench.py --workload array --steps 3 --warmup 3 ) 2>&1 | tail -1 >> $L
NGB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 0 > gpurun_out/final_ncu.log 2>&1
NGB_NO_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:ngb_k_lu_packed -s 30 -c 1 -f -o gpurun_out/final_lu python tests/gpu_profile_run.py 4096 > gpurun_out/final_lu.log 2>&1
grep -v "^$" $L | cut -c1-1200 | tail -40
