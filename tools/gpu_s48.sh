#!/bin/bash
# session 48: k_neural residency / walk-round variants under the class-ordered drain
mkdir -p gpurun_out/s48
for v in t4c1 t5c1 t6c1 wr1 wr3; do
  lib=$PWD/nbvh_b200/libnbvh_b200_$v.so; [ $v = default ] && lib=$PWD/nbvh_b200/libnbvh_b200.so
  NBVH_B200_LIB=$lib timeout 300 python tools/perf_r2.py --frame-only > gpurun_out/s48/frame_$v.json 2> gpurun_out/s48/frame_$v.err
  echo "== $v"; cut -c1-520 gpurun_out/s48/frame_$v.json; tail -2 gpurun_out/s48/frame_$v.err | cut -c1-200
done
