#!/bin/bash
mkdir -p gpurun_out/s70
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/s70/pytest.log
timeout 300 python tools/perf_r2.py --frame-only > gpurun_out/s70/frame.json 2> gpurun_out/s70/frame.err; cut -c1-300 gpurun_out/s70/frame.json
