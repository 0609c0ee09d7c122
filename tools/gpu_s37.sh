#!/bin/bash
# slab test with FMNMX / FMNMX3 on the hoisted-division path: parity suite + frame / path / sampler timings
TAG=${1:-s37}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 $OUT/${TAG}_pytest.log | cut -c1-300
timeout 600 python tools/perf_r2.py ${TAG}_perf.json > $OUT/${TAG}_perf.log 2>&1; echo "perf rc=$?"; head -1 $OUT/${TAG}_perf.log | cut -c1-600; python - <<PY
import json
d=json.load(open("$OUT/${TAG}_perf.json")) if __import__("os").path.exists("$OUT/${TAG}_perf.json") else {}
print({k:d.get(k) for k in ("sampler","train_2p18")})
PY
timeout 300 python tools/path_perf.py ${TAG}_path.json > $OUT/${TAG}_path.log 2>&1; tail -3 $OUT/${TAG}_path.log | cut -c1-300
