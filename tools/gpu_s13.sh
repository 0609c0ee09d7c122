#!/bin/bash
# streaming frames: bit-identity test under a watchdog, perf with and without
TAG=${1:-s13}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_stream_frames.py -q -m gpu -x > $OUT/${TAG}_pytest_stream.log 2>&1; echo "stream pytest rc=$?"
tail -15 $OUT/${TAG}_pytest_stream.log | cut -c1-300
NBVH_STREAM_FRAMES=1 timeout 300 python tools/perf_r2.py ${TAG}_perf_stream.json --quick --frame-only > $OUT/${TAG}_perf_stream.log 2>&1; echo "perf(stream) rc=$?"; head -1 $OUT/${TAG}_perf_stream.log | cut -c1-900
timeout 300 python tools/perf_r2.py ${TAG}_perf.json --quick --frame-only > $OUT/${TAG}_perf.log 2>&1; echo "perf rc=$?"; head -1 $OUT/${TAG}_perf.log | cut -c1-900
timeout 600 python -m pytest tests/test_gpu_training.py tests/test_tcnn_goldens.py -q -m gpu > $OUT/${TAG}_pytest_train.log 2>&1; echo "train pytest rc=$?"; tail -4 $OUT/${TAG}_pytest_train.log | cut -c1-300
timeout 600 python tools/perf_r2.py ${TAG}_perf_full.json --quick > $OUT/${TAG}_perf_full.log 2>&1; echo "perf rc=$?"; tail -12 $OUT/${TAG}_perf_full.log
