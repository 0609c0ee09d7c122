#!/bin/bash
# slim k_train_mlp epilogues: training + sampler tests, perf, ncu --set full of the optimisation kernels
TAG=${1:-s12}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_sampler.py tests/test_gpu_training.py -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $OUT/${TAG}_pytest.log | cut -c1-400
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_training.py -q -m gpu -k long_loss -s 2>&1 | grep -E "median loss|passed|failed" | tr '\n' ' '; echo; done
timeout 600 python tools/perf_r2.py ${TAG}_perf.json --quick > $OUT/${TAG}_perf.log 2>&1; echo "perf rc=$?"; tail -12 $OUT/${TAG}_perf.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_train_mlp|k_grid_backward|k_encode_features|k_adam' --launch-skip 10 -c 5 -f -o $OUT/${TAG}_train \
    python tools/ncu_driver.py > $OUT/${TAG}_ncu_train.log 2>&1
tail -2 $OUT/${TAG}_ncu_train.log
