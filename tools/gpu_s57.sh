#!/bin/bash
mkdir -p gpurun_out/s57
for v in sp4 sp8 sp16 default; do
  lib=$PWD/nbvh_b200/libnbvh_b200_$v.so; [ $v = default ] && lib=$PWD/nbvh_b200/libnbvh_b200.so
  NBVH_B200_LIB=$lib timeout 300 python tools/perf_r2.py --frame-only > gpurun_out/s57/frame_$v.json 2> gpurun_out/s57/frame_$v.err
  echo "== $v"; cut -c1-480 gpurun_out/s57/frame_$v.json; tail -1 gpurun_out/s57/frame_$v.err | cut -c1-200
done
