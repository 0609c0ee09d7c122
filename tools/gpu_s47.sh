#!/bin/bash
# session 47: five-class drain order as default (camera rays of both pipelines): GPU suite, frame timing, bench (both arms)
TAG=s47
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log | cut -c1-200
timeout 300 python tools/perf_r2.py --frame-only > $OUT/${TAG}_frame.json 2> $OUT/${TAG}_frame.err; cut -c1-500 $OUT/${TAG}_frame.json
NBVH_CHAIN_CLASSES=0 timeout 300 python tools/perf_r2.py --frame-only > $OUT/${TAG}_frame_off.json 2> $OUT/${TAG}_frame_off.err; cut -c1-500 $OUT/${TAG}_frame_off.json
timeout 900 python bench.py --impl reference > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "bench ref rc=$?"
timeout 1200 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-1500 $OUT/${TAG}_bench.json
NBVH_CHAIN_CLASSES=0 timeout 1200 python bench.py --no-reference-gpu > $OUT/${TAG}_bench_off.json 2> $OUT/${TAG}_bench_off.err; echo "bench(off) rc=$?"; cut -c1-1500 $OUT/${TAG}_bench_off.json
