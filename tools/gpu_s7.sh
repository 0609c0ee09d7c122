#!/bin/bash
# k_neural look-ahead: full GPU tests (bit-exactness of frames / hit records), perf of default + variants
TAG=${1:-s7}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 $OUT/${TAG}_pytest.log
timeout 600 python tools/perf_r2.py ${TAG}_perf.json > $OUT/${TAG}_perf.log 2>&1; echo "perf rc=$?"; head -1 $OUT/${TAG}_perf.log | cut -c1-900; tail -12 $OUT/${TAG}_perf.log
for v in nbvh_b200/libnbvh_b200_*.so; do
  [ -f "$v" ] || continue
  n=$(basename $v .so | sed 's/libnbvh_b200_//')
  NBVH_B200_LIB=$PWD/$v timeout 300 python tools/perf_r2.py ${TAG}_perf_$n.json --quick > $OUT/${TAG}_perf_$n.log 2>&1
  echo "== variant $n rc=$?"; head -1 $OUT/${TAG}_perf_$n.log | cut -c1-1300
done
