#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/exp9.log; : > $L
echo "== bench default (grouped)" >> $L; ( timeout 1500 python bench.py --steps 1 --warmup 3 ) 2>&1 | tail -1 >> $L
echo "== bench unsorted" >> $L; ( NGB_BENCH_UNSORTED=1 timeout 1500 python bench.py --steps 1 --warmup 3 ) 2>&1 | tail -1 >> $L
python - <<'PY'
import json
for ln in open("gpurun_out/exp9.log"):
    if ln.startswith("{"):
        d=json.loads(ln); print(d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"])
    else: print(ln.strip())
PY
