#!/bin/bash
mkdir -p gpurun_out/s56
for v in pf1 pf2 pf3 default; do
  lib=$PWD/nbvh_b200/libnbvh_b200_$v.so; [ $v = default ] && lib=$PWD/nbvh_b200/libnbvh_b200.so
  NBVH_B200_LIB=$lib timeout 300 python tools/perf_r2.py --frame-only > gpurun_out/s56/frame_$v.json 2> gpurun_out/s56/frame_$v.err
  echo "== $v"; cut -c1-480 gpurun_out/s56/frame_$v.json
  NBVH_B200_LIB=$lib timeout 300 python tools/path_perf.py s56/path_$v.json 2>&1 | cut -c1-120
done
