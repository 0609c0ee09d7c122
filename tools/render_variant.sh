#!/bin/bash
# Development: libnbvh_b200_<name>.so = the default objects with render.cu recompiled under extra -D switches
# usage: tools/render_variant.sh <name> -DNBVH_X=1 ...
set -e
name=$1; shift
cd "$(dirname "$0")/.."
O=nbvh_b200/_obj/default
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC,-O2,-ffp-contract=off "$@" -Xptxas -v -c nbvh_b200/csrc/render.cu -o /tmp/render_$name.o 2>&1 | grep -A2 "k_neuralILb0ELb0" | grep "spill\|Used"
objs=$(ls $O/*.o | grep -v "/render.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o nbvh_b200/libnbvh_b200_$name.so $objs /tmp/render_$name.o -ldl
