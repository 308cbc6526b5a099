#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/exp7.log; : > $L
for i in 1 2; do
echo "== sweep, branches on" >> $L; ( timeout 600 python bench.py --workload sweep --steps 3 --warmup 3 ) 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])" >> $L
echo "== sweep, branches off" >> $L; ( NGB_NO_BRANCH=1 timeout 600 python bench.py --workload sweep --steps 3 --warmup 3 ) 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])" >> $L
done
echo "== ro101" >> $L; ( timeout 600 python bench.py --workload ro101 --steps 2 --warmup 3 ) 2>&1 | tail -1 | cut -c1-400 >> $L
cat $L
