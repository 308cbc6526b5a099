#!/bin/bash
# gpu_r2_grid.sh: the grid-wide LU on the 16 x 16 array (parity + time per Newton step), then the usual suite
mkdir -p gpurun_out; L=gpurun_out/r2_grid.log; : > $L
( timeout 600 python -m pytest tests/test_synth_array.py -m gpu -x -q -s 2>&1 | tail -6 ) >> $L
( timeout 300 python - <<'PY'
import sys, time; sys.path.insert(0, 'tests')
from parity_util import *
lib = pkg.Library()
flat = ngt.read(f"{GOLDEN}/arr16.flat.ngt.gz"); pats = ngt.read(f"{GOLDEN}/arr16.pat.ngt.gz"); wave = ngt.read(f"{GOLDEN}/arr16.wave.ngt")
circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=run_patterns(pats))
b = pkg.Batch(circ, 1)
for rep in range(2):
    t0 = time.time(); res = b.tran(1024, wave["save_eq"]); dt = time.time() - t0
    print("arr16 tran", dt, "s", int(res.numiter[0]), "iterations", dt / int(res.numiter[0]) * 1e6, "us per Newton step")
PY
) >> $L 2>&1
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) >> $L
cat $L
