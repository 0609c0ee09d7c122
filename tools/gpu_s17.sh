#!/bin/bash
TAG=${1:-s17}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_training.py tests/test_tcnn_goldens.py tests/test_tree_cut.py -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $OUT/${TAG}_pytest.log | cut -c1-300
NBVH_B200_LIB=$PWD/nbvh_b200/libnbvh_b200_traintiming.so timeout 200 python tools/train_timing_probe.py 2>&1 | tail -3
timeout 600 python tools/perf_r2.py ${TAG}_perf.json --quick > $OUT/${TAG}_perf.log 2>&1; echo "perf rc=$?"; tail -12 $OUT/${TAG}_perf.log
