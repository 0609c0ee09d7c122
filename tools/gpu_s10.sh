#!/bin/bash
# k_train_mlp with two tiles in flight: GPU suite (training parity), perf with one and two slots
TAG=${1:-s10}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 $OUT/${TAG}_pytest.log | cut -c1-400
timeout 600 python tools/perf_r2.py ${TAG}_perf.json --quick > $OUT/${TAG}_perf.log 2>&1; echo "perf rc=$?"; tail -22 $OUT/${TAG}_perf.log
NBVH_TRAIN_SLOTS=1 timeout 600 python tools/perf_r2.py ${TAG}_perf_slots1.json --quick > $OUT/${TAG}_perf_slots1.log 2>&1; echo "perf(1 slot) rc=$?"; tail -12 $OUT/${TAG}_perf_slots1.log
