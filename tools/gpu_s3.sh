#!/bin/bash
# Round-2 GPU session 3 (1 GPU): full GPU tests, the new bench.py (both arms, live reference-GPU arm), launch list.
TAG=${1:-s3}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
( time timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
tail -5 $OUT/${TAG}_bench.err
python - <<'PY'
import json,sys
try:
    j=json.load(open('gpurun_out/'+sys.argv[1]+'_bench.json')) if len(sys.argv)>1 else None
except Exception as e:
    print('no bench json', e); j=None
PY
python -c "
import json
j=json.load(open('$OUT/${TAG}_bench.json'))
print(json.dumps({k:j[k] for k in ('value','ms_per_step','e2e','roofline','config')}, indent=1)[:6000])
print(json.dumps(j.get('train'), indent=1)[:4000])
print(json.dumps(j.get('reference_gpu_b200'), indent=1)[:2500])
print(json.dumps(j.get('path'), indent=1)[:2500])
"
timeout 600 python tools/perf_r2.py ${TAG}_perf.json --quick > $OUT/${TAG}_perf.log 2>&1; echo "perf rc=$?"; tail -45 $OUT/${TAG}_perf.log
