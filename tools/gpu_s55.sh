#!/bin/bash
mkdir -p gpurun_out/s55
for v in sparse4 sparse1; do
NBVH_B200_LIB=$PWD/nbvh_b200/libnbvh_b200_$v.so timeout 300 python tools/perf_r2.py --frame-only > gpurun_out/s55/frame_$v.json 2> gpurun_out/s55/frame_$v.err
python - <<PY
import json
for l in open('gpurun_out/s55/frame_$v.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print("$v", json.dumps(d.get("k_neural_phase_cycles_per_iteration")), d["k_neural_timeline_us"])
PY
grep "probe\] mlp" gpurun_out/s55/frame_$v.json | tail -1
done
