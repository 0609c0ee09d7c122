#!/bin/bash
mkdir -p gpurun_out/s58
timeout 900 python -m pytest tests/test_path.py tests/test_ref_goldens.py tests/test_ray_batches.py -q -m gpu -x 2>&1 | tail -5
for pr in 1 9; do
  NBVH_CHAIN_PRED=$pr timeout 300 python tools/perf_r2.py --frame-only > gpurun_out/s58/frame_p$pr.json 2> gpurun_out/s58/frame_p$pr.err
  echo "== pred flags $pr"; cut -c1-480 gpurun_out/s58/frame_p$pr.json; tail -1 gpurun_out/s58/frame_p$pr.err | cut -c1-200
  NBVH_CHAIN_PRED=$pr timeout 300 python tools/path_perf.py s58/path_p$pr.json 2>&1 | cut -c1-120
done
