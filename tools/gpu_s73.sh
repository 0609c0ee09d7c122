#!/bin/bash
mkdir -p gpurun_out/s73
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/s73/pytest.log
timeout 600 python bench.py > gpurun_out/s73/bench.json 2> gpurun_out/s73/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/s73/bench.json
