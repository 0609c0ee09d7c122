#!/bin/bash
# k_neural: converged MMA issuer (default build) vs previous library; walk budget, tile-cooperative and chunk-major gather variants.
TAG=${1:-s38}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $OUT/${TAG}_pytest.log | cut -c1-300
timeout 600 python tools/perf_r2.py ${TAG}_perf.json --frame-only > $OUT/${TAG}_perf.log 2>&1; echo "perf rc=$?"; head -1 $OUT/${TAG}_perf.log | cut -c1-500
timeout 300 python tools/path_perf.py ${TAG}_path.json > $OUT/${TAG}_path.log 2>&1; tail -3 $OUT/${TAG}_path.log | cut -c1-300
for v in nbvh_b200/libnbvh_b200_*.so; do
  [ -f "$v" ] || continue
  n=$(basename $v .so | sed 's/libnbvh_b200_//')
  NBVH_B200_LIB=$PWD/$v timeout 300 python tools/perf_r2.py ${TAG}_perf_$n.json --frame-only > $OUT/${TAG}_perf_$n.log 2>&1
  echo "== variant $n rc=$?"; head -1 $OUT/${TAG}_perf_$n.log | cut -c1-500
  NBVH_B200_LIB=$PWD/$v timeout 300 python tools/path_perf.py ${TAG}_path_$n.json > $OUT/${TAG}_path_$n.log 2>&1; tail -3 $OUT/${TAG}_path_$n.log | cut -c1-300
done
