#!/bin/bash
# round-2 baseline on HEAD: GPU parity tests, FP64 peak, Newton-step timing, launch list, full ncu capture of the load kernel
mkdir -p gpurun_out; L=gpurun_out/r2_base.log; : > $L
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -4 >> $L
python - >> $L 2>&1 <<'PY'
import ctypes, sys
sys.path.insert(0, "tests")
from parity_util import pkg
lib = pkg.library(); lib.check(lib.L.ngbInit(0), "init")
out = (ctypes.c_double * 3)()
print("fp64peak rc", lib.L.ngbMeasureFp64Peak(out), "dfma TFLOP/s %.2f  dadd/dmul TFLOP/s %.2f  clock kHz %.0f" % (out[0] / 1e12, out[1] / 1e12, out[2]))
PY
cd tests
for i in 1 2; do ( timeout 120 python gpu_profile_run.py 4096 ) 2>&1 | tail -1 >> ../$L; done
cd ..
NGB_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_base_launches.csv python tests/gpu_profile_run.py 4096 > gpurun_out/r2_base_ncu.log 2>&1
NGB_NO_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:ngb_k_bsim4_load -s 30 -c 1 -f -o gpurun_out/r2_base_b4load python tests/gpu_profile_run.py 4096 > gpurun_out/r2_base_b4load.log 2>&1
ncu -i gpurun_out/r2_base_b4load.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_base_b4load_source.csv 2>/dev/null
ls -la gpurun_out >> $L
cat $L
