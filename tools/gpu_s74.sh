#!/bin/bash
mkdir -p gpurun_out/s74
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/memcheck_queue_order.py > gpurun_out/s74/memcheck.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/s74/memcheck.log | cut -c1-200
