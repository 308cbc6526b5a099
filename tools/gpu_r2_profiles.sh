#!/bin/bash
# round-2 evidence on HEAD: GPU suite, launch list of the default bench, full ncu captures of the three kernels of a Newton
# step, the default bench line and the bench lines of the workloads the driver does not run (sweep, ro101, array, array_tran)
mkdir -p gpurun_out; L=gpurun_out/r2_prof.log; : > $L
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) >> $L
NGB_NO_GRAPH=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 0 > gpurun_out/r2_prof_ncu1.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_bench.csv >> $L 2>&1
NGB_NO_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:ngb_k_bsim4_load -s 30 -c 1 -f -o gpurun_out/r02_b4load python tests/gpu_profile_run.py 4096 > gpurun_out/r2_prof_ncu2.log 2>&1
NGB_NO_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:ngb_k_lu_packed -s 30 -c 1 -f -o gpurun_out/r02_lu python tests/gpu_profile_run.py 4096 > gpurun_out/r2_prof_ncu3.log 2>&1
NGB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none -k regex:ngb_k_assemble -s 30 -c 1 -f -o gpurun_out/r02_asm python tests/gpu_profile_run.py 4096 > gpurun_out/r2_prof_ncu4.log 2>&1
( cd tests; timeout 120 python gpu_profile_run.py 4096 stages 2>&1 | tail -2 ) >> $L
( time timeout 1200 python bench.py ) > gpurun_out/r02_bench_default.json 2>> $L
tail -c 400 gpurun_out/r02_bench_default.json >> $L
for w in ${WORKLOADS:-sweep ro101 array}; do
  ( time timeout 900 python bench.py --workload $w ) > gpurun_out/r02_bench_$w.json 2>> $L
  tail -c 600 gpurun_out/r02_bench_$w.json >> $L
done
( timeout 600 python bench.py --workload array_tran --cells 2304 --steps 1 --warmup 3 2>&1 | tail -1 ) > gpurun_out/r02_bench_array_tran_2304.json
tail -c 600 gpurun_out/r02_bench_array_tran_2304.json >> $L
ls -la gpurun_out | tail -20 >> $L
cat $L
