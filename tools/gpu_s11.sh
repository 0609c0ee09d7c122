#!/bin/bash
# sampler options (view-dependent with the golden's camera, bounce diagnosis), k_train_mlp issuer with hoisted descriptors, two-issuer variant
TAG=${1:-s11}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_sampler.py tests/test_gpu_training.py -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 $OUT/${TAG}_pytest.log | cut -c1-400
timeout 600 python tools/perf_r2.py ${TAG}_perf.json --quick > $OUT/${TAG}_perf.log 2>&1; echo "perf rc=$?"; tail -12 $OUT/${TAG}_perf.log
NBVH_B200_LIB=$PWD/nbvh_b200/libnbvh_b200_twoiss.so timeout 600 python -m pytest tests/test_gpu_training.py -q -m gpu > $OUT/${TAG}_pytest_twoiss.log 2>&1; echo "pytest(two issuers) rc=$?"; tail -3 $OUT/${TAG}_pytest_twoiss.log | cut -c1-300
NBVH_B200_LIB=$PWD/nbvh_b200/libnbvh_b200_twoiss.so timeout 600 python tools/perf_r2.py ${TAG}_perf_twoiss.json --quick > $OUT/${TAG}_perf_twoiss.log 2>&1; echo "perf(two issuers) rc=$?"; tail -12 $OUT/${TAG}_perf_twoiss.log
