#!/bin/bash
# option goldens from the reference kernels (sampler options, WorldT), full GPU suite, perf of default + variants
TAG=${1:-s9}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/make_ref_sampler_option_goldens.py > $OUT/${TAG}_optgold.log 2>&1; echo "option goldens rc=$?"; grep "option goldens\]" $OUT/${TAG}_optgold.log | cut -c1-200; tail -3 $OUT/${TAG}_optgold.log | cut -c1-400
timeout 300 python tools/make_ref_worldt_golden.py > $OUT/${TAG}_worldt.log 2>&1; echo "worldt golden rc=$?"; tail -2 $OUT/${TAG}_worldt.log | cut -c1-300
cp -f $OUT/ref_golden/ref_chess_sampler_options.npz $OUT/ref_golden/ref_chess_worldt_256x144.npz tests/golden/ 2>/dev/null
timeout 1200 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $OUT/${TAG}_pytest.log | cut -c1-300
timeout 600 python tools/perf_r2.py ${TAG}_perf.json > $OUT/${TAG}_perf.log 2>&1; echo "perf rc=$?"; head -1 $OUT/${TAG}_perf.log | cut -c1-900; tail -6 $OUT/${TAG}_perf.log
for v in nbvh_b200/libnbvh_b200_*.so; do
  [ -f "$v" ] || continue
  n=$(basename $v .so | sed 's/libnbvh_b200_//')
  NBVH_B200_LIB=$PWD/$v timeout 300 python tools/perf_r2.py ${TAG}_perf_$n.json --quick > $OUT/${TAG}_perf_$n.log 2>&1
  echo "== variant $n rc=$?"; head -1 $OUT/${TAG}_perf_$n.log | cut -c1-1300
done
