#!/bin/bash
# session 45: drain order by expected chain length, class-limit sweep (NBVH_CHAIN_CLASSES) x predictor (NBVH_CHAIN_PRED)
mkdir -p gpurun_out/s45
run() { # pred classes
  f=gpurun_out/s45/frame_p$1_c$(echo $2 | tr ',' '_').json
  NBVH_CHAIN_PRED=$1 NBVH_CHAIN_CLASSES=$2 timeout 300 python tools/perf_r2.py --frame-only > $f 2> ${f%.json}.err
  echo "pred=$1 classes=$2"; cut -c1-420 $f
}
run 1 0
run 1 4,4,4,4
run 1 6,4,3,2
run 0 6,4,3,2
run 3 6,4,3,2
run 1 8,5,3,2
run 1 5,3,2,2
run 1 6,4,2,2
run 1 10,6,4,3
run 1 7,5,4,3
run 1 4,3,2,2
