#!/bin/bash
# oracle/flops_gcov.sh -- TEST INFRASTRUCTURE ONLY.  Pins F_alg, the algorithmic flop count of one BSIM4
# instance-evaluation: compiles the reference's b4ld.c with --coverage (-O0) where it lies,
# links it in front of oracle/_ref/libngref.a, runs one Monte-Carlo sample of config 3 (ro17k.cir) and
# feeds gcov's per-line execution counts to tools/flops_gcov.py.  Needs oracle/build_ref.sh to have run.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
R="$REF/src"
G="$HERE/_ref/gcov"
mkdir -p "$G" && cd "$G" && rm -f b4ld.gcda
gcc -c -O0 --coverage -w -fPIC -fno-strict-aliasing -I"$HERE/ref_config" -I"$R/include" -I"$R/spicelib/devices" -I"$R/maths/KLU" \
    -I"$R/frontend" -I"$R/spicelib/parser" -I"$R/maths/poly" -I"$R/maths/sparse" -I"$R" -I"$R/spicelib/devices/bsim4" \
    "$R/spicelib/devices/bsim4/b4ld.c" -o b4ld.o
gcc --coverage -o ngspice_gcov ../obj_serial/main.o b4ld.o -Wl,--start-group ../libngref.a -Wl,--end-group -lm -ldl
./ngspice_gcov -b -r "$G/run.raw" "$HERE/../tests/golden/netlists/ro17k.cir" > run.log 2>&1
rm -f "$G/run.raw"
gcov -o . "$R/spicelib/devices/bsim4/b4ld.c" > gcov.log 2>&1
python3 "$HERE/../tools/flops_gcov.py" "$G/b4ld.c.gcov"
