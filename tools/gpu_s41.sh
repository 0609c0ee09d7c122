#!/bin/bash
# config 5 sweep on the final build + per-kernel launch list of one config-2b frame
TAG=${1:-s41}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python bench.py --sweep > $OUT/${TAG}_sweep.json 2> $OUT/${TAG}_sweep.err; echo "sweep rc=$?"; cut -c1-600 $OUT/${TAG}_sweep.json
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file $OUT/${TAG}_path_launches.csv python tools/path_ncu.py 1920 1080 3 > $OUT/${TAG}_path_ncu.log 2>&1; echo "path ncu rc=$?"
python - <<PY
import csv
rows=list(csv.reader(open("$OUT/${TAG}_path_launches.csv")))
hdr=None; out=[]
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); out.append((d["ID"],d["Kernel Name"][:40],d["Metric Name"][:40],d["Metric Value"]))
for o in out[-60:]: print(o)
PY
