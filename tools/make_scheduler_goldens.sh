#!/bin/bash
# Goldens of the reference's BvhSplitScheduler (host integer logic) -> tests/golden/split_scheduler.json
set -e
REF=/root/reference
mkdir -p oracle/_ref
g++ -std=c++17 -I$REF/include -I$REF/ext/glm -I$REF/ext/tiny-cuda-nn/dependencies -DGLM_ENABLE_EXPERIMENTAL \
    oracle/ref_tools/split_scheduler_golden.cpp -o oracle/_ref/split_scheduler_golden
{
  echo '{"cases": ['
  echo '{"args": [40, 3000, 1, 2.5, 2.0, 10000], "out":'; oracle/_ref/split_scheduler_golden 40 3000 1 2.5 2.0 10000; echo '},'
  echo '{"args": [50, 2000, 1, 2.0, 2.0, 4000], "out":'; oracle/_ref/split_scheduler_golden 50 2000 1 2.0 2.0 4000; echo '},'
  echo '{"args": [30, 900, 3, 1.7, 1.5, 1500], "out":'; oracle/_ref/split_scheduler_golden 30 900 3 1.7 1.5 1500; echo '}'
  echo ']}'
} > tests/golden/split_scheduler.json
python -c "import json; j=json.load(open('tests/golden/split_scheduler.json')); print([c['out']['stats'] for c in j['cases']])"
