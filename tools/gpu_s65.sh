#!/bin/bash
mkdir -p gpurun_out/s65
NBVH_ENCODE_MERGE=1 timeout 900 python -m pytest tests/test_gpu_training.py -q -m gpu -x 2>&1 | tail -3
for m in 1 0 1 0; do
NBVH_ENCODE_MERGE=$m timeout 900 python bench.py --no-reference-gpu > gpurun_out/s65/bench_m$m.json 2> gpurun_out/s65/bench_m$m.err
python - <<PY
import json
d=json.load(open('gpurun_out/s65/bench_m$m.json'))
r=d['roofline']; print("encode_merge=$m encode_ms", round(r['k_encode_features_ms'],4), "step", round(d['train']['ms_per_step'],4), round(d['train']['value']/1e6,1))
PY
done
