#!/bin/bash
# GPU session: reference-kernel goldens + reference frame time, tests, bench (both arms), ncu launch list and full captures.
# usage (under gpurun): bash tools/gpu_round2.sh <tag>
TAG=${1:-r01g}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt
if [ -x oracle/_ref/ref_render_harness ]; then
  timeout 900 python tools/make_ref_render_goldens.py > $OUT/${TAG}_goldens.log 2>&1; echo "goldens rc=$?" | tee -a $OUT/${TAG}_goldens.log
  tail -8 $OUT/${TAG}_goldens.log
  cp -f $OUT/ref_golden/ref_chess_256x144.npz $OUT/ref_golden/ref_chess_1080p.npz $OUT/ref_golden/ref_chess_timing.json tests/golden/ 2>/dev/null
fi
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench_ref.json
timeout 600 python bench.py --steps 200 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
cat $OUT/${TAG}_bench.json
timeout 300 python tools/gpu_perf.py > $OUT/${TAG}_perf.json 2>> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_perf.json
for v in nbvh_b200/libnbvh_b200_*.so; do
  [ -f "$v" ] || continue
  n=$(basename $v .so | sed 's/libnbvh_b200_//')
  NBVH_B200_LIB=$PWD/$v timeout 300 python tools/gpu_perf.py > $OUT/${TAG}_perf_$n.json 2>> $OUT/${TAG}_bench.err
  echo "variant $n"; cat $OUT/${TAG}_perf_$n.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_neural|k_primary' -s 6 -c 4 -f -o $OUT/${TAG}_frame \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_query -s 30 -c 1 -f -o $OUT/${TAG}_query \
    python tools/gpu_perf.py > $OUT/${TAG}_ncu_query.log 2>&1
ls -la $OUT
