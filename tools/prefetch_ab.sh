#!/bin/bash
# A/B of the per-frame L2 warm-up: bench headline (L2 flushed between frames) with and without it
for v in 1 0; do
  echo "== NBVH_L2_PREFETCH=$v"
  NBVH_L2_PREFETCH=$v python bench.py --steps 100 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.1f ms %.4f e2e %.1f k_neural %.4f k_primary %.4f path2b %.3f path4 %.2f' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['kernel_ms'], j['roofline_primary']['kernel_ms'], j['path']['config2b']['ms_per_frame'], j['path']['config4']['ms_per_16spp_image']))"
done
