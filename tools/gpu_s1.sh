#!/bin/bash
# Round-2 GPU session 1: full GPU test suite, perf of the default build and the A/B variants, live reference-GPU arm, ncu launch
# list + full captures (with source) of the frame kernels, the two-pass sample generator and every kernel of the optimisation step.
TAG=${1:-s1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt
timeout 900 python -m pytest tests -q -m gpu -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -12 $OUT/${TAG}_pytest.log
timeout 600 python tools/perf_r2.py ${TAG}_perf.json > $OUT/${TAG}_perf.log 2>&1; echo "perf rc=$?"; tail -70 $OUT/${TAG}_perf.log
for v in nbvh_b200/libnbvh_b200_*.so; do
  [ -f "$v" ] || continue
  n=$(basename $v .so | sed 's/libnbvh_b200_//')
  NBVH_B200_LIB=$PWD/$v timeout 300 python tools/perf_r2.py ${TAG}_perf_$n.json --quick > $OUT/${TAG}_perf_$n.log 2>&1
  echo "== variant $n rc=$?"; head -1 $OUT/${TAG}_perf_$n.log | cut -c1-700
done
timeout 600 python tools/ref_gpu_arm.py > $OUT/${TAG}_ref_gpu_arm.json 2> $OUT/${TAG}_ref_gpu_arm.err; echo "ref arm rc=$?"; tail -3 $OUT/${TAG}_ref_gpu_arm.err; head -c 1500 $OUT/${TAG}_ref_gpu_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv \
    python tools/ncu_driver.py > $OUT/${TAG}_ncu_list.log 2>&1
# full captures: the LAST frame and the LAST optimisation step of the driver (skip the warm-up launches by kernel-name count)
NCU_WARM=1 timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'k_neural|k_primary|k_sample_walk|k_sample_targets|k_encode_features|k_train_mlp|k_grid_backward|k_adam' \
    --launch-skip 4 -c 9 -f -o $OUT/${TAG}_full python tools/ncu_driver.py > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_ncu_full.log
ls -la $OUT | head -40
