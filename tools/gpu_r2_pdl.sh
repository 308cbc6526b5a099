#!/bin/bash
# programmatic dependent launch (NGB_PDL): GPU suite without it, Newton-step timing at levels 0 / 1 / 2, GPU suite with level 2
mkdir -p gpurun_out; L=gpurun_out/r2_pdl.log; : > $L
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) >> $L
cd tests
for lvl in 0 1 2 0 2; do
  echo "== NGB_PDL=$lvl" >> ../$L
  for i in 1 2; do ( NGB_PDL=$lvl timeout 120 python gpu_profile_run.py 4096 ) 2>&1 | tail -1 >> ../$L; done
done
cd ..
echo "== suite with NGB_PDL=2" >> $L
( NGB_PDL=2 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) >> $L
cat $L
