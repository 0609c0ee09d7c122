#!/bin/bash
mkdir -p gpurun_out/s61
timeout 900 python bench.py --sweep > gpurun_out/s61/sweep.json 2> gpurun_out/s61/sweep.err; echo "sweep rc=$?"; cut -c1-900 gpurun_out/s61/sweep.json
