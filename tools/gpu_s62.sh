#!/bin/bash
mkdir -p gpurun_out/s62
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_shipped_configs.py -q -m gpu -x 2>&1 | tail -3
for m in 0 1; do
NBVH_GRID_BACKWARD_PER_POINT=$m timeout 900 python bench.py --no-reference-gpu > gpurun_out/s62/bench_pp$m.json 2> gpurun_out/s62/bench_pp$m.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/s62/bench_pp$m.json'))
r=d['roofline']; print("per_point=$m", {k:round(r[k],4) for k in r if k.endswith('_ms')}); print(d['train']['ms_per_step'], d['train']['value'])
PY
done
