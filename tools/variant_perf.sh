#!/bin/bash
# A/B timing of the k_neural build variants (tools: nbvh_b200/build.py --variant): frame times of configs 2a / 2b / 4
OUT=gpurun_out
TAG=${1:-var}
python tools/path_perf.py ${TAG}_base.json > $OUT/${TAG}_base.log 2>&1; tail -3 $OUT/${TAG}_base.log
for v in nbvh_b200/libnbvh_b200_*.so; do
  n=$(basename $v .so | sed 's/libnbvh_b200_//')
  NBVH_B200_LIB=$PWD/$v timeout 300 python tools/path_perf.py ${TAG}_$n.json > $OUT/${TAG}_$n.log 2>&1
  echo "== variant $n"; tail -3 $OUT/${TAG}_$n.log | cut -c1-200
done
