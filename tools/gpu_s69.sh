#!/bin/bash
mkdir -p gpurun_out/s69
for v in pt64 pt128 pt512 default; do
  lib=$PWD/nbvh_b200/libnbvh_b200_$v.so; [ $v = default ] && lib=$PWD/nbvh_b200/libnbvh_b200.so
  NBVH_B200_LIB=$lib timeout 300 python tools/perf_r2.py --frame-only > gpurun_out/s69/frame_$v.json 2> gpurun_out/s69/frame_$v.err
  echo "== $v"; cut -c1-260 gpurun_out/s69/frame_$v.json
done
