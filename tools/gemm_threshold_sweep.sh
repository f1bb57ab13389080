#!/bin/bash
# which bra blocks should go to the DMMA kernel?  (RMB_GEMM_MIN_DK: blocks with dim_k above it)
for t in 12 9 7 5 3; do
  RMB_GEMM_MIN_DK=$t python bench.py --no-cpu-baseline --steps 30 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('min_dk', $t, round(d['value']), 'state-steps/s, matvec avg us', round(r['avg_launch_us'],1), 'TF', round(r['fp64']['achieved'],2))"
done
