#!/bin/bash
mkdir -p gpurun_out/s76
timeout 90 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_training.py -q -m gpu -k "forward_backward_matches or partial_tile" > gpurun_out/s76/memcheck_train.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/s76/memcheck_train.log | cut -c1-200
