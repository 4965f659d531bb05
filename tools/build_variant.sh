#!/bin/bash
# Builds an experimental variant of the library: tools/build_variant.sh <name> [-DFLAG ...] -> build/variants/lib_<name>.so
# (objects are cached under build/obj; conv_tc is recompiled when its sources or the variant's flags changed)
set -e
NAME=$1; shift
FL="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
C=codon_b200/csrc
mkdir -p build/obj build/variants
newest_hdr=$(ls -t $C/*.h $C/*.cuh include/*.h | head -1)
for f in conv_direct cac_nchw edge metrics preproc; do
  if [ ! -f build/obj/$f.o ] || [ $C/$f.cu -nt build/obj/$f.o ] || [ $newest_hdr -nt build/obj/$f.o ]; then nvcc $FL -c -o build/obj/$f.o $C/$f.cu; fi
done
# api.cu, conv_tc.cu and cac.cu take the variant's flags (-DCODON_TC_EXPERIMENT enables the environment knobs)
for f in api conv_tc cac; do
  O=build/obj/${f}_$NAME.o
  if [ ! -f $O ] || [ $C/$f.cu -nt $O ] || [ $newest_hdr -nt $O ] || [ "$(cat $O.flags 2>/dev/null)" != "$*" ]; then
    nvcc $FL "$@" -c -o $O $C/$f.cu 2>&1 | grep -E "error" || true
    echo "$*" > $O.flags
  fi
done
nvcc -shared -o build/variants/lib_$NAME.so build/obj/{conv_direct,cac_nchw,edge,metrics,preproc}.o build/obj/api_$NAME.o build/obj/conv_tc_$NAME.o build/obj/cac_$NAME.o
ls -la build/variants/lib_$NAME.so
