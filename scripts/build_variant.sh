#!/bin/bash
# Experimental builds of libmargin_head.so with extra -D flags (A/B measurements on one box):
#   scripts/build_variant.sh n256 -DMH_S_TILE_N=256   ->  face_recognition_models_b200/libmargin_head_n256.so
# select it with MH_LIB=/root/repo/face_recognition_models_b200/libmargin_head_n256.so
set -e
name=$1; shift
src=$(cd "$(dirname "$0")/../face_recognition_models_b200/csrc" && pwd)
out=/tmp/mh_variant_$name; mkdir -p $out
for f in capi prologue dense stash verify tc_head step; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $src/$f.cu -o $out/$f.o 2>/dev/null &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $src/../libmargin_head_$name.so $out/*.o -lcudart
echo built libmargin_head_$name.so
