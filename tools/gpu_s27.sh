#!/bin/bash
# flat persistent trace kernel for bounce rays / ray batches: full GPU suite, path perf A/B (NBVH_TRACE_FLAT=0 = thread per path)
TAG=${1:-s27}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 $OUT/${TAG}_pytest.log | cut -c1-300
timeout 300 python tools/path_perf.py ${TAG}_path.json > $OUT/${TAG}_path.log 2>&1; tail -3 $OUT/${TAG}_path.log | cut -c1-300
NBVH_TRACE_FLAT=0 timeout 300 python tools/path_perf.py ${TAG}_path_tpp.json > $OUT/${TAG}_path_tpp.log 2>&1; echo "== thread per path"; tail -3 $OUT/${TAG}_path_tpp.log | cut -c1-300
timeout 600 python bench.py --sweep --steps 5 --warmup 2 > $OUT/${TAG}_sweep.json 2> $OUT/${TAG}_sweep.err; echo "sweep rc=$?"; cut -c1-1500 $OUT/${TAG}_sweep.json
NBVH_TRACE_FLAT=0 timeout 600 python bench.py --sweep --steps 5 --warmup 2 > $OUT/${TAG}_sweep_tpp.json 2> $OUT/${TAG}_sweep_tpp.err; echo "sweep tpp rc=$?"; cut -c1-1500 $OUT/${TAG}_sweep_tpp.json
