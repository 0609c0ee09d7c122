#!/bin/bash
# A/B timing of build variants (nbvh_b200/build.py --variant): frame times of configs 2a / 2b / 4 + network query rate
OUT=gpurun_out
TAG=${1:-var}
python tools/path_perf.py ${TAG}_base.json > $OUT/${TAG}_base.log 2>&1; tail -3 $OUT/${TAG}_base.log | cut -c1-120
python tools/gpu_perf.py 2>/dev/null | grep -A3 "query_1048576\|render_1080p" | grep "ms\"" 
for v in nbvh_b200/libnbvh_b200_*.so; do
  [ -f "$v" ] || continue
  n=$(basename $v .so | sed 's/libnbvh_b200_//')
  NBVH_B200_LIB=$PWD/$v timeout 300 python tools/path_perf.py ${TAG}_$n.json > $OUT/${TAG}_$n.log 2>&1
  echo "== variant $n"; tail -3 $OUT/${TAG}_$n.log | cut -c1-120
  NBVH_B200_LIB=$PWD/$v python tools/gpu_perf.py 2>/dev/null | grep -A3 "query_1048576\|render_1080p" | grep "ms\""
done
