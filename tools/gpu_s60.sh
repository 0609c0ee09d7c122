#!/bin/bash
mkdir -p gpurun_out/s60
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_tree_cut.py -q -m gpu -x 2>&1 | tail -3
timeout 900 python bench.py --no-reference-gpu > gpurun_out/s60/bench.json 2> gpurun_out/s60/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s60/bench.json'))
r=d['roofline']; print(d['value'], {k:r[k] for k in r if k.startswith('k_') or k=='sample_generation_ms'}); print(d['train']['ms_per_step'], d['train']['value'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_reduce_wgrad|k_train_mlp' -c 12 --csv python tools/ncu_driver.py 2>/dev/null | grep -i "k_reduce\|k_train" | cut -d, -f5,15- | tail -8
