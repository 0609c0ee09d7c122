#!/bin/bash
# Round-2 GPU session 2: GPU tests (sampler v3 with cooperative tail), perf of default + k_neural variants incl. phase clocks.
TAG=${1:-s2}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -12 $OUT/${TAG}_pytest.log
timeout 600 python tools/perf_r2.py ${TAG}_perf.json --quick > $OUT/${TAG}_perf.log 2>&1; echo "perf rc=$?"; tail -60 $OUT/${TAG}_perf.log
for v in nbvh_b200/libnbvh_b200_*.so; do
  [ -f "$v" ] || continue
  n=$(basename $v .so | sed 's/libnbvh_b200_//')
  NBVH_B200_LIB=$PWD/$v timeout 300 python tools/perf_r2.py ${TAG}_perf_$n.json --quick > $OUT/${TAG}_perf_$n.log 2>&1
  echo "== variant $n rc=$?"; head -1 $OUT/${TAG}_perf_$n.log | cut -c1-1200
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sample_walk|k_sample_targets' --launch-skip 2 -c 2 -f -o $OUT/${TAG}_sampler \
    python tools/ncu_driver.py > $OUT/${TAG}_ncu_sampler.log 2>&1
tail -2 $OUT/${TAG}_ncu_sampler.log
