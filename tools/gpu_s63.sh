#!/bin/bash
mkdir -p gpurun_out/s63
for m in 0 1 2 3 6; do
NBVH_GRID_BACKWARD_MERGE_LEVELS=$m timeout 900 python bench.py --no-reference-gpu > gpurun_out/s63/bench_m$m.json 2> gpurun_out/s63/bench_m$m.err
python - <<PY
import json
d=json.load(open('gpurun_out/s63/bench_m$m.json'))
r=d['roofline']; print("merge_levels=$m grid_backward_ms", round(r['k_grid_backward_ms'],4), "step", round(d['train']['ms_per_step'],4), round(d['train']['value']/1e6,1))
PY
done
