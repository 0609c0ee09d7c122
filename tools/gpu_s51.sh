#!/bin/bash
# session 51: two-frame history maximum (NBVH_CHAIN_PRED bit 2)
mkdir -p gpurun_out/s51
for pr in 5 1 5 1; do
  NBVH_CHAIN_PRED=$pr timeout 300 python tools/perf_r2.py --frame-only > gpurun_out/s51/frame_p$pr.json 2> gpurun_out/s51/frame_p$pr.err
  echo "== pred $pr"; cut -c1-470 gpurun_out/s51/frame_p$pr.json
done
NBVH_CHAIN_PRED=5 NBVH_CHAIN_CLASSES=5,4,3,2 timeout 300 python tools/perf_r2.py --frame-only > gpurun_out/s51/frame_p5_c5432.json; echo "== pred 5 classes 5,4,3,2";  cut -c1-470 gpurun_out/s51/frame_p5_c5432.json
NBVH_CHAIN_PRED=5 NBVH_B200_LIB=$PWD/nbvh_b200/libnbvh_b200_cstats.so timeout 300 python tools/chain_stats.py 6 > gpurun_out/s51/chain_stats_p5.txt 2>&1
tail -17 gpurun_out/s51/chain_stats_p5.txt | head -10
