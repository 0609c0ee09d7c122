#!/bin/bash
# Round-2 GPU session 4 (N GPUs, default 2): in-library data-parallel checks, bench.py at N, multi-GPU C++ train host.
N=${1:-2}
TAG=${2:-s4}
OUT=gpurun_out
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $RUN tools/dp_grad_check.py > $OUT/${TAG}_dp_grad.json 2> $OUT/${TAG}_dp_grad.err; echo "dp_grad rc=$?"; cat $OUT/${TAG}_dp_grad.json; tail -3 $OUT/${TAG}_dp_grad.err
timeout 300 $RUN tools/dp_tree_cut_check.py 120 > $OUT/${TAG}_dp_treecut.json 2> $OUT/${TAG}_dp_treecut.err; echo "dp_treecut rc=$?"; cat $OUT/${TAG}_dp_treecut.json; tail -3 $OUT/${TAG}_dp_treecut.err
( time timeout 900 $RUN bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err ) 2>&1 | grep real; echo "bench rc=$?"
tail -5 $OUT/${TAG}_bench_${N}gpu.err
python -c "
import json
j=json.load(open('$OUT/${TAG}_bench_${N}gpu.json'))
print(json.dumps({k:j.get(k) for k in ('value','ms_per_step','n_gpus','e2e','config','strong_scaling_one_frame')}, indent=1)[:5000])
t=j.get('train') or {}
print(json.dumps({k:t.get(k) for k in ('value','ms_per_step','strong_scaling','sample_generation_ms')}, indent=1)[:2500])
print(json.dumps(t.get('roofline'), indent=1)[:2500])
"
rm -rf /tmp/cli_out; mkdir -p /tmp/cli_out
NBVH_NCCL_ID_FILE=/tmp/cli_out/nccl_id timeout 600 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  nbvh_b200/host/nbvh_train_cli baseline/_ref/chess nodes_32940_hash_18_lvl_8_tsteps_10000 /tmp/cli_out 300 2 640 360 > $OUT/${TAG}_train_cli.log 2>&1; echo "train_cli rc=$?"
grep "^{" $OUT/${TAG}_train_cli.log; tail -3 $OUT/${TAG}_train_cli.log | cut -c1-300
timeout 300 nbvh_b200/host/nbvh_train_cli baseline/_ref/chess nodes_32940_hash_18_lvl_8_tsteps_10000 /tmp/cli_out1 300 2 640 360 2>&1 | grep "^{"
