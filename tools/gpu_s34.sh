#!/bin/bash
TAG=${1:-s34}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/make_ref_uniform_golden.py > $OUT/${TAG}_uniform_golden.log 2>&1; echo "golden rc=$?"; tail -4 $OUT/${TAG}_uniform_golden.log | cut -c1-300
NBVH_B200_LIB=$PWD/nbvh_b200/libnbvh_b200_phase.so timeout 600 python tools/perf_r2.py ${TAG}_phase.json --frame-only > $OUT/${TAG}_phase.log 2>&1; echo "phase rc=$?"; grep -v "^\[k_neural probe\]" $OUT/${TAG}_phase.log | head -1 | cut -c1-1500; grep "probe\] mlp" $OUT/${TAG}_phase.log | tail -1; grep "probe\] walk" $OUT/${TAG}_phase.log | tail -1
