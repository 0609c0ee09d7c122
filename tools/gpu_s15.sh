#!/bin/bash
# event-driven issuer + split commits: full GPU suite, perf, 1-GPU bench
TAG=${1:-s15}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 $OUT/${TAG}_pytest.log | cut -c1-300
timeout 600 python tools/perf_r2.py ${TAG}_perf.json --quick > $OUT/${TAG}_perf.log 2>&1; echo "perf rc=$?"; tail -12 $OUT/${TAG}_perf.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/s15_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e'].get('synchronous_value'))
print(json.dumps(d['train']['roofline'])[:1500])
print(d['config'].get('config2b_ms_per_step'), d['config'].get('config4_16spp_image_ms'), d['config'].get('train_ms_per_step'))
P
