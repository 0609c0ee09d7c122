#!/bin/bash
# Round-2 session 30 (N GPUs, default 8): bench.py at N exactly as the driver launches it.
N=${1:-8}
TAG=${2:-s30}
OUT=gpurun_out
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time timeout 900 $RUN bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err ) 2>&1 | grep real; echo "bench rc=$?"
tail -5 $OUT/${TAG}_bench_${N}gpu.err
python -c "
import json
j=json.load(open('$OUT/${TAG}_bench_${N}gpu.json'))
print(json.dumps({k:j.get(k) for k in ('value','ms_per_step','n_gpus','e2e','config')}, indent=1)[:4000])
t=j.get('train') or {}
print(json.dumps({k:t.get(k) for k in ('value','ms_per_step','strong_scaling','sample_generation_ms')}, indent=1)[:2500])
"
