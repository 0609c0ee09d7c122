#!/bin/bash
# session 49: lane-pair gather variant (NBVH_PAIR_GATHER) -- frame timing + parity tests under the variant library
mkdir -p gpurun_out/s49
for v in pair default; do
  lib=$PWD/nbvh_b200/libnbvh_b200_$v.so; [ $v = default ] && lib=$PWD/nbvh_b200/libnbvh_b200.so
  NBVH_B200_LIB=$lib timeout 300 python tools/perf_r2.py --frame-only > gpurun_out/s49/frame_$v.json 2> gpurun_out/s49/frame_$v.err
  echo "== $v"; cut -c1-520 gpurun_out/s49/frame_$v.json; tail -2 gpurun_out/s49/frame_$v.err | cut -c1-200
  NBVH_B200_LIB=$lib timeout 300 python tools/path_perf.py s49/path_$v.json 2>&1 | cut -c1-200
done
NBVH_B200_LIB=$PWD/nbvh_b200/libnbvh_b200_pair.so timeout 900 python -m pytest tests/test_ref_goldens.py tests/test_path.py -q -m gpu -x 2>&1 | tail -3
