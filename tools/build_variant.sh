#!/bin/bash
# build_variant.sh NAME [nvcc defines...] -- builds build/variants/NAME/libngb200.so (experiment builds for
# one GPU call that compares several kernels; selected with NGB200_LIB=...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=build/variants/$name
mkdir -p $out
PKG=ngspice-sf-mirror_b200; CSRC=$PKG/csrc
[ -f $CSRC/ngb_host.o ] || make -s all
python3 tools/nvcc_outline.py --outline-entries bsim4,ngb_k_b4_ --inline-div ${NGB_INLINE_DIV:-none} --share-rcp ${NGB_SHARE_RCP:-0} -- nvcc \
  -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -Xcompiler -fPIC -std=c++17 -Xptxas -v \
  -I$CSRC -Iinclude "$@" -c $CSRC/ngb_cuda.cu -o $out/ngb_cuda.o 2> $out/ptxas.log || { cat $out/ptxas.log; exit 1; }
nvcc -shared -o $out/libngb200.so $out/ngb_cuda.o $CSRC/ngb_host.o $CSRC/ngb_tran.o $CSRC/ngb_pivot.o $CSRC/ngb_b4temp.o -lcudart -lgomp
grep -A2 "ngb_k_bsim4_load\|ngb_k_lu_packed" $out/ptxas.log | grep -v "^--" | grep "registers\|spill" | tr '\n' ' '; echo
