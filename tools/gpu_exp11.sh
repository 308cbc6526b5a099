#!/bin/bash
for i in 1 2 3; do ( timeout 600 python bench.py --workload sweep --steps 4 --warmup 3 ) 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sweep', d['value'], d['ms_per_step'], d['e2e']['value'])"; done
cd tests; timeout 120 python gpu_profile_run.py 4096 | tail -1
