#!/bin/bash
# Round-2 final measurements on one B200: full GPU suite, bench.py (both arms), ncu launch list of the bench command, full captures
# (with source) of the frame kernels, the sample generator and every kernel of the optimisation step.
TAG=${1:-s39}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log | cut -c1-200
timeout 900 python bench.py --impl reference > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "bench ref rc=$?"
timeout 1200 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-900 $OUT/${TAG}_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-reference-gpu > $OUT/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
NCU_WARM=1 timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'k_neural|k_primary|k_sample_walk|k_sample_targets|k_node_keep|k_encode_features|k_train_mlp|k_reduce_wgrad|k_grid_backward|k_adam' \
    --launch-skip 4 -c 14 -f -o $OUT/${TAG}_full python tools/ncu_driver.py > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_ncu_full.log
ls -la $OUT/${TAG}_*
