#!/bin/bash
mkdir -p gpurun_out/s75
timeout 110 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_path.py -q -m gpu -k "queue_drain_order" > gpurun_out/s75/memcheck_path.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/s75/memcheck_path.log | cut -c1-200
