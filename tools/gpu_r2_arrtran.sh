#!/bin/bash
# gpu_r2_arrtran.sh [cells...]: config 4 with the LU (DC op + transient, grid-wide LU) at the given sizes (default 48x48, 64x64,
# 128x128)
mkdir -p gpurun_out; L=gpurun_out/r2_arrtran.log; : > $L
( timeout 600 python -m pytest tests/test_synth_array.py -m gpu -x -q -s 2>&1 | tail -6 ) >> $L
for c in ${@:-2304 4096 16384}; do
  echo "== cells $c" >> $L
  ( timeout 900 python bench.py --workload array_tran --cells $c --steps 1 --warmup 3 2>&1 | tail -1 ) > gpurun_out/r02_bench_array_tran_$c.json
  python - $c >> $L <<'PY'
import json, sys
c = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r02_bench_array_tran_{c}.json").read().strip().splitlines()[-1])
    print("ms/step", round(d["ms_per_step"], 1), "iters", d["newton_iterations"], "us/iter", round(d["us_per_newton_iteration"], 1), "stages", {k: round(v, 1) for k, v in d["stage_us_per_sampled_step"].items()})
    print("  cpu us/iter", round(d["cpu_baseline"]["us_per_newton_iteration"], 1), "analysis_s", d["cpu_baseline"]["analysis_s"], "parity", d["parity_check"]["ok"], d["parity_check"]["max_rel_err"])
except Exception as e:
    print("ERR", e, open(f"gpurun_out/r02_bench_array_tran_{c}.json").read()[-500:])
PY
done
cat $L
