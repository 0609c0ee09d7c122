#!/bin/bash
# Region split of depth-1 frames (NBVH_REGION_SPLIT): parity suite + frame timings with the split on and off.
TAG=${1:-s36}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 $OUT/${TAG}_pytest.log | cut -c1-300
for m in 1 0; do
  NBVH_REGION_SPLIT=$m timeout 600 python tools/perf_r2.py ${TAG}_perf_split$m.json --frame-only > $OUT/${TAG}_perf_split$m.log 2>&1; echo "perf split=$m rc=$?"; head -1 $OUT/${TAG}_perf_split$m.log | cut -c1-600
done
NBVH_REGION_SPLIT=1 timeout 600 python bench.py --steps 100 --warmup 5 --no-reference-gpu > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-400 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
