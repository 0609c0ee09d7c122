bash tools/gpu_s39.sh s66
