#!/bin/bash
# Round-end GPU session: full GPU test suite, both bench arms, launch list of the bench, full ncu captures of the path kernels.
TAG=${1:-r02f}
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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_neural|k_primary' -s 6 -c 4 -f -o $OUT/${TAG}_frame \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_path_trace|k_path_shade|k_neural' -s 18 -c 9 -f -o $OUT/${TAG}_path \
    python tools/path_ncu.py 1920 1080 3 > $OUT/${TAG}_ncu_path.log 2>&1
ls -la $OUT | grep ${TAG}
