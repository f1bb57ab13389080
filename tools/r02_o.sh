#!/bin/bash
out=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $out/r02o_tests.log 2>&1
tail -12 $out/r02o_tests.log
for wl in ocs_align ocs_mixed; do
for fl in 0 1; do
RMB_FUSED_LIN=$fl timeout 600 python bench.py --workload $wl --steps 2000 --no-cpu-baseline --also none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl fused_lin=$fl value', round(d['value'],1), 'us/step', round(d['ms_per_step']*1e3,1), 'e2e', round(d['e2e']['value'],1), 'parity', d['parity']['ok'], d['parity']['parity_max_rel'])"
done
done
