#!/bin/bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
# Compiles the reference ngspice sources WHERE THEY LIE under /root/reference
# (no copy) with plain gcc into oracle/_ref/ (git-ignored).  autotools/bison are
# not available, so this replaces the reference build system:
#   - ref_config/ngspice/config.h      hand-written config.h (KLU on, XSPICE/CIDER/OSDI off)
#   - ref_config/{parse-bison,inpptree-parser}.h + ref_pp_parser.c/ref_pt_parser.c
#                                      hand-written stand-ins for the two bison grammars
# Produces:
#   oracle/_ref/libngref.a     every reference object (serial flavour, -O2)
#   oracle/_ref/ngspice        stock reference CLI binary      (the CPU oracle / CPU baseline)
#   oracle/_ref/ngspice_omp    same with USE_OMP (-fopenmp)    (OpenMP CPU baseline), if OMP=1
# Usage: oracle/build_ref.sh [REF=/root/reference]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
if [ ! -d "$REF/src" ]; then echo "build_ref: $REF/src not present - skipping (prebuilt oracle/_ref is used)"; exit 0; fi
R="$REF/src"
OUT="$HERE/_ref"
FLAV="${FLAV:-serial}"
OBJ="$OUT/obj_$FLAV"
mkdir -p "$OBJ"
CFLAGS="-O2 -w -fPIC -fno-strict-aliasing"
[ "$FLAV" = omp ] && CFLAGS="$CFLAGS -fopenmp -DUSE_OMP"
INC="-I$HERE/ref_config -I$R/include -I$R/spicelib/devices -I$R/maths/KLU -I$R/frontend -I$R/spicelib/parser -I$R/maths/poly -I$R/maths/sparse -I$R"
export R OBJ CFLAGS INC

list() {
  find "$R/spicelib" "$R/maths" "$R/misc" "$R/frontend" -name '*.c' \
    | grep -v -E '/maths/KLU/|/maths/dense/|/test_|/devices/(adms|nbjt|nbjt2|numd|numd2|numos|ndev|hicum2)/|/frontend/(wdisp|help)/|/plotting/(x11|xgraph)|/maths/sparse/spsmp\.c|/parser/inp2n\.c|/analysis/(dcpss|pssinit|psssetp|pssaskq|psssetparm)\.c|/frontend/testcommands\.c|/frontend/nutmegif\.c|/misc/(getopt_long_bsd|tilde_win)\.c|/maths/fft/fftext_test' 
  echo "$R/main.c"; echo "$R/conf.c"; echo "$R/ngspice.c"
}
cc1() {
  f="$1"; o="$OBJ/$(echo "${f#$R/}" | sed 's#/#_#g; s#\.c$#.o#')"
  [ "$o" -nt "$f" ] && return 0
  X=""; case "$f" in "$R/main.c"|"$R/conf.c"|"$R/ngspice.c") X="-DSIMULATOR";; esac
  gcc -c $CFLAGS $X $INC -I"$(dirname "$f")" "$f" -o "$o" 2>"$o.err" || { echo "FAIL $f"; cat "$o.err" | head -5; return 1; }
  rm -f "$o.err"
}
export -f cc1
list | xargs -P "$(nproc)" -n 1 bash -c 'cc1 "$0"'

# KLU: numeric files twice (real, complex), the rest once (src/maths/KLU/Makefile.am)
K="$R/maths/KLU"
for f in klu klu_diagnostics klu_dump klu_extract klu_factor klu_free_numeric klu_kernel klu_multiply klu_refactor klu_scale klu_solve klu_sort klu_tsolve klu_utils; do
  [ "$OBJ/klur_$f.o" -nt "$K/$f.c" ] || gcc -c $CFLAGS $INC "$K/$f.c" -o "$OBJ/klur_$f.o" &
  [ "$OBJ/kluz_$f.o" -nt "$K/$f.c" ] || gcc -c $CFLAGS $INC -DCOMPLEX "$K/$f.c" -o "$OBJ/kluz_$f.o" &
done
for f in amd_1 amd_2 amd_aat amd_control amd_defaults amd_dump amd_global amd_info amd_order amd_postorder amd_post_tree amd_preprocess amd_valid btf_maxtrans btf_order btf_strongcomp colamd colamd_global klu_analyze klu_analyze_given klu_defaults klu_free_symbolic klu_memory klusmp; do
  [ "$OBJ/klu1_$f.o" -nt "$K/$f.c" ] || gcc -c $CFLAGS $INC "$K/$f.c" -o "$OBJ/klu1_$f.o" &
done
wait
# hand-written stand-ins for the bison outputs
gcc -c $CFLAGS $INC -I"$R/frontend" "$HERE/ref_pp_parser.c" -o "$OBJ/x_ref_pp_parser.o"
gcc -c $CFLAGS $INC -I"$R/spicelib/parser" "$HERE/ref_pt_parser.c" -o "$OBJ/x_ref_pt_parser.o"
gcc -c $CFLAGS $INC "$HERE/ref_stubs.c" -o "$OBJ/x_ref_stubs.o"

SUF=""; [ "$FLAV" = omp ] && SUF="_omp"
rm -f "$OUT/libngref$SUF.a"
MAINOBJ="$OBJ/main.o"
ar rcs "$OUT/libngref$SUF.a" $(ls "$OBJ"/*.o | grep -v "/main.o$")
LDOMP=""; [ "$FLAV" = omp ] && LDOMP="-fopenmp"
gcc $LDOMP -o "$OUT/ngspice$SUF" "$MAINOBJ" -Wl,--start-group "$OUT/libngref$SUF.a" -Wl,--end-group -lm -ldl
echo "built $OUT/ngspice$SUF"
if [ "$FLAV" = serial ]; then
  # instrumented flavour: same objects + oracle/ref_hooks.c interposed with ld --wrap
  gcc -c $CFLAGS $INC -I"$R/spicelib/devices" -I"$R/maths/KLU" "$HERE/ref_hooks.c" -o "$OBJ/y_ref_hooks.o.tmp" && mv "$OBJ/y_ref_hooks.o.tmp" "$OUT/ref_hooks.o"
  gcc -o "$OUT/ngspice_dump" "$MAINOBJ" "$OUT/ref_hooks.o" \
      -Wl,--wrap=CKTload,--wrap=SMPsolve,--wrap=SMPluFac,--wrap=SMPreorder,--wrap=DCtran \
      -Wl,--start-group "$OUT/libngref.a" -Wl,--end-group -lm -ldl
  echo "built $OUT/ngspice_dump"
  # drop-in flavour: the same objects with integration/ngb_shim.c interposed in front of CKTload and
  # the SMP entry points, linked against the product library (INTEGRATION.md level 1); a second copy is
  # linked against the host build of the kernels so the binding can be exercised without a GPU
  REPO="$(cd "$HERE/.." && pwd)"
  gcc -c $CFLAGS $INC -I"$R/spicelib/devices" -I"$R/maths/KLU" "$REPO/integration/ngb_shim.c" -o "$OUT/ngb_shim.o"
  if [ -f "$REPO/ngspice-sf-mirror_b200/libngb200.so" ]; then
    gcc -o "$OUT/ngspice_ngb" "$MAINOBJ" "$OUT/ngb_shim.o" \
        -Wl,--wrap=CKTload,--wrap=SMPsolve,--wrap=SMPluFac,--wrap=SMPreorder \
        -Wl,--start-group "$OUT/libngref.a" -Wl,--end-group -L"$REPO/ngspice-sf-mirror_b200" -lngb200 \
        -Wl,-rpath,'$ORIGIN/../../ngspice-sf-mirror_b200' -Wl,-rpath,/usr/local/cuda/lib64 -lm -ldl && echo "built $OUT/ngspice_ngb"
  fi
  if [ -f "$REPO/tests/hostsim/libngb200_hostsim.so" ]; then
    gcc -o "$OUT/ngspice_ngb_hostsim" "$MAINOBJ" "$OUT/ngb_shim.o" \
        -Wl,--wrap=CKTload,--wrap=SMPsolve,--wrap=SMPluFac,--wrap=SMPreorder \
        -Wl,--start-group "$OUT/libngref.a" -Wl,--end-group -L"$REPO/tests/hostsim" -lngb200_hostsim \
        -Wl,-rpath,'$ORIGIN/../../tests/hostsim' -lm -ldl -lstdc++ && echo "built $OUT/ngspice_ngb_hostsim"
  fi
fi
