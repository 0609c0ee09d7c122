#!/bin/bash
# sampler iteration: sampler + training tests, perf (quick), ncu of the sampler kernels
TAG=${1:-s5}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_sampler.py tests/test_gpu_training.py tests/test_tree_cut.py tests/test_edge_cases.py -q -m gpu -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 $OUT/${TAG}_pytest.log
timeout 600 python tools/perf_r2.py ${TAG}_perf.json --quick > $OUT/${TAG}_perf.log 2>&1; echo "perf rc=$?"; tail -42 $OUT/${TAG}_perf.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sample_walk|k_sample_targets' --launch-skip 3 -c 3 -f -o $OUT/${TAG}_sampler \
    python tools/ncu_driver.py > $OUT/${TAG}_ncu_sampler.log 2>&1
tail -2 $OUT/${TAG}_ncu_sampler.log
