#!/bin/bash
# overlay reading of the per-sample parameter rows + __grid_constant__ kernel parameters: GPU suite, the lock-step timing batch
# (shared rows: shows the __grid_constant__ effect alone), the default bench with the overlay and with NGB_B4_OVERLAY=0
mkdir -p gpurun_out; L=gpurun_out/r2_overlay.log; : > $L
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) >> $L
( cd tests; for i in 1 2; do timeout 120 python gpu_profile_run.py 4096 2>&1 | tail -1; done; timeout 120 python gpu_profile_run.py 4096 stages 2>&1 | tail -1 ) >> $L
echo "== bench, overlay" >> $L
( timeout 600 python bench.py --steps 1 --warmup 1 2>&1 | tail -1 ) > gpurun_out/r2_overlay_on.json; python tools/bench_brief.py gpurun_out/r2_overlay_on.json >> $L
echo "== bench, NGB_B4_OVERLAY=0" >> $L
( NGB_B4_OVERLAY=0 timeout 600 python bench.py --steps 1 --warmup 1 2>&1 | tail -1 ) > gpurun_out/r2_overlay_off.json; python tools/bench_brief.py gpurun_out/r2_overlay_off.json >> $L
cat $L
