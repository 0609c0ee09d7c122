#!/bin/bash
# session 43: chain-length history (NBVH_LONG_CHAIN) A/B on the headline frame + full GPU test suite of the rebuilt tree
mkdir -p gpurun_out/s43
for t in 0 4 3 5 6 8 2; do
  NBVH_LONG_CHAIN=$t timeout 300 python tools/perf_r2.py --frame-only > gpurun_out/s43/frame_lc$t.json 2> gpurun_out/s43/frame_lc$t.err
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s43/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s43/pytest_gpu.log
tail -3 gpurun_out/s43/pytest_gpu.log
for t in 0 4 3 5 6 8 2; do echo "lc=$t"; cat gpurun_out/s43/frame_lc$t.json; done
