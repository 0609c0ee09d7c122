#!/bin/bash
# parity at bench scale: 4K reference golden (+ measured metrics), then the full GPU suite with the new tests
TAG=${1:-s8}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python tools/make_ref_path4k_golden.py > $OUT/${TAG}_golden4k.log 2>&1; echo "golden rc=$?"; tail -4 $OUT/${TAG}_golden4k.log | cut -c1-400
cp -f $OUT/ref_golden/ref_chess_path_4k_16spp.npz tests/golden/ 2>/dev/null
timeout 1200 python -m pytest tests -q -m gpu -s > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "logit-space|config 4 \(4K|passed|failed|FAILED|Error" $OUT/${TAG}_pytest.log | head -20
