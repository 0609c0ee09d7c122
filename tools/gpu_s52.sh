#!/bin/bash
mkdir -p gpurun_out/s52
NBVH_B200_LIB=$PWD/nbvh_b200/libnbvh_b200_phase.so timeout 300 python tools/perf_r2.py --frame-only > gpurun_out/s52/frame_phase.json 2> gpurun_out/s52/frame_phase.err
python - <<'PY'
import json
for l in open('gpurun_out/s52/frame_phase.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps({k:d[k] for k in d if k!='frame_8spl'},indent=1))
PY
grep "probe" gpurun_out/s52/frame_phase.json | tail -2
