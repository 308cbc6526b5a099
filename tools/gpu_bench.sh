#!/bin/bash
# gpu_bench.sh TAG [bench.py arguments]: GPU parity tests (unless SKIP_TESTS), then one bench.py line, both logged under gpurun_out/
TAG=$1; shift
mkdir -p gpurun_out; L=gpurun_out/$TAG.log; : > $L
if [ -z "$SKIP_TESTS" ]; then ( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) >> $L; fi
( time timeout 1500 python bench.py "$@" ) > gpurun_out/${TAG}_bench.json 2>> $L
tail -1 gpurun_out/${TAG}_bench.json >> $L
cat $L
