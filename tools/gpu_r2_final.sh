#!/bin/bash
# evidence of the round's final build: launch list of the default bench (first 400 launches), `ncu --set full` capture of the
# BSIM4 load kernel at Newton step 15 000 of the bench's transient (the regime the bench averages over), the default bench line
mkdir -p gpurun_out; L=gpurun_out/r2_final.log; : > $L
NGB_NO_GRAPH=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --kill 1 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 0 > gpurun_out/r2_final_ncu1.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_bench.csv >> $L 2>&1
NGB_NO_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:ngb_k_bsim4_load -s 15000 -c 1 --kill 1 -f \
  -o gpurun_out/r02_b4load python bench.py --steps 1 --warmup 1 > gpurun_out/r2_final_ncu2.log 2>&1
ncu -i gpurun_out/r02_b4load.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_b4load_source.csv 2>/dev/null
ncu -i gpurun_out/r02_b4load.ncu-rep --page details > gpurun_out/r02_b4load_ncu_details.txt 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02_b4load.ncu-rep > gpurun_out/r02_b4load_ncu_key_metrics.txt 2>/dev/null
tail -2 gpurun_out/r2_final_ncu2.log >> $L
( time timeout 1200 python bench.py ) > gpurun_out/r02_bench_default.json 2>> $L
python tools/bench_brief.py gpurun_out/r02_bench_default.json >> $L
cat $L
