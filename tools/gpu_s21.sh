#!/bin/bash
# distance-skipped walk-stack entries: full GPU suite (bit-exact interval lists / hit records), perf A/B (NBVH_NO_STACK_SKIP=1), probe build
TAG=${1:-s21}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 $OUT/${TAG}_pytest.log | cut -c1-300
timeout 600 python tools/perf_r2.py ${TAG}_perf.json > $OUT/${TAG}_perf.log 2>&1; echo "perf rc=$?"; head -1 $OUT/${TAG}_perf.log | cut -c1-900
NBVH_NO_STACK_SKIP=1 timeout 600 python tools/perf_r2.py ${TAG}_perf_noskip.json --frame-only > $OUT/${TAG}_perf_noskip.log 2>&1; echo "noskip rc=$?"; head -1 $OUT/${TAG}_perf_noskip.log | cut -c1-600
NBVH_B200_LIB=$PWD/nbvh_b200/libnbvh_b200_phase.so timeout 300 python tools/perf_r2.py --frame-only 2>&1 | grep "probe\]" | tail -4
