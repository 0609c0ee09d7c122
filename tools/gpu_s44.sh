#!/bin/bash
# session 44: chain-length history predictor variants (NBVH_CHAIN_PRED bit0 = 3x3 max, bit1 = decay) x thresholds
mkdir -p gpurun_out/s44
for pr in 1 3 2; do for t in 3 4 5 6; do
  NBVH_CHAIN_PRED=$pr NBVH_LONG_CHAIN=$t timeout 300 python tools/perf_r2.py --frame-only > gpurun_out/s44/frame_p${pr}_lc$t.json 2> gpurun_out/s44/frame_p${pr}_lc$t.err
  echo "pred=$pr lc=$t"; cut -c1-330 gpurun_out/s44/frame_p${pr}_lc$t.json
done; done
