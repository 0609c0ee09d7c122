#!/bin/bash
# GPU session for the path-tracing rows: reference path-tracer goldens + timings, GPU tests, path timings.
# usage (under gpurun): bash tools/gpu_round3.sh <tag> [nogold] [full]
TAG=${1:-r01s}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt
if [ -x oracle/_ref/ref_render_harness ] && [ "$2" != "nogold" ]; then
  timeout 900 python tools/make_ref_path_goldens.py > $OUT/${TAG}_path_goldens.log 2>&1; echo "path goldens rc=$?" | tee -a $OUT/${TAG}_path_goldens.log
  tail -12 $OUT/${TAG}_path_goldens.log
  cp -f $OUT/ref_golden/ref_chess_path_256x144.npz $OUT/ref_golden/ref_chess_path_timing.json tests/golden/ 2>/dev/null
fi
timeout 900 python -m pytest tests/test_path.py -q -m gpu -s > $OUT/${TAG}_pytest_path.log 2>&1; echo "pytest path rc=$?" | tee -a $OUT/${TAG}_pytest_path.log
tail -25 $OUT/${TAG}_pytest_path.log
if [ "$3" == "full" ]; then
  timeout 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_path.py > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
  tail -8 $OUT/${TAG}_pytest.log
fi
timeout 600 python tools/path_perf.py ${TAG}_path_perf.json > $OUT/${TAG}_path_perf.log 2>&1; echo "perf rc=$?"
cat $OUT/${TAG}_path_perf.log | tail -8
