#!/bin/bash
mkdir -p gpurun_out/s72
for v in sh128 sh64 default; do
  lib=$PWD/nbvh_b200/libnbvh_b200_$v.so; [ $v = default ] && lib=$PWD/nbvh_b200/libnbvh_b200.so
  echo "== $v"; NBVH_B200_LIB=$lib timeout 300 python tools/path_perf.py s72/path_$v.json 2>&1 | cut -c1-110
done
