#!/bin/bash
# Lean round-end GPU session (no kernel changed since r02q/r02r): full GPU test suite, smoke, both bench arms, launch list, C++ host run.
TAG=${1:-r02t}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -6 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
head -c 400 $OUT/${TAG}_bench.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
nbvh_b200/host/nbvh_render_cli baseline/_ref/chess nodes_32940_hash_18_lvl_8_tsteps_10000 /tmp/cli.ppm 1920 1080 64 1 > $OUT/${TAG}_cpp_host.json 2>&1
cat $OUT/${TAG}_cpp_host.json
ls -la $OUT | grep ${TAG}
