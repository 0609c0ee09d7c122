#!/bin/bash
mkdir -p gpurun_out/s46
NBVH_B200_LIB=$PWD/nbvh_b200/libnbvh_b200_cstats.so timeout 300 python tools/chain_stats.py 4 > gpurun_out/s46/chain_stats.txt 2>&1
tail -17 gpurun_out/s46/chain_stats.txt
