#!/bin/bash
# Builds an experimental variant of the library: tools/build_variant.sh <name> [-DFLAG ...] -> build/variants/lib_<name>.so
set -e
NAME=$1; shift
FL="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
C=codon_b200/csrc
for f in api conv_direct cac cac_nchw edge metrics preproc; do
  if [ ! -f build/obj/$f.o ] || [ $C/$f.cu -nt build/obj/$f.o ]; then nvcc $FL -c -o build/obj/$f.o $C/$f.cu; fi
done
nvcc $FL "$@" -c -o build/obj/conv_tc_$NAME.o $C/conv_tc.cu 2>&1 | grep -E "error" || true
nvcc -shared -o build/variants/lib_$NAME.so build/obj/{api,conv_direct,cac,cac_nchw,edge,metrics,preproc}.o build/obj/conv_tc_$NAME.o
ls -la build/variants/lib_$NAME.so
