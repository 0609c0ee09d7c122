#!/bin/bash
# Uniform segment points on the wavefront path: golden from the reference's kernels, then the parity tests (CPU leg + GPU leg).
TAG=${1:-s35}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/make_ref_uniform_golden.py > $OUT/${TAG}_uniform_golden.log 2>&1; echo "golden rc=$?"; tail -3 $OUT/${TAG}_uniform_golden.log | cut -c1-300
cp $OUT/ref_golden/ref_chess_uniform_256x144.npz tests/golden/ 2>/dev/null
timeout 900 python -m pytest tests/test_ref_goldens.py tests/test_path.py -q -x -k "uniform or worldt or hit_records or shading_branches" > $OUT/${TAG}_pytest_uniform.log 2>&1; echo "pytest uniform rc=$?"
tail -15 $OUT/${TAG}_pytest_uniform.log | cut -c1-300
timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $OUT/${TAG}_pytest.log | cut -c1-300
