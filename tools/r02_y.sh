#!/bin/bash
# k_matvec_lin instruction diet: probe, soak (both tiles, single-group variant), lin / ocs tests, bench line
out=gpurun_out
timeout 300 python tools/matvec_probe.py ocs_batch 8192 100 2>&1 | tail -2
timeout 600 python tools/lin_soak.py 300 512 T8,T4,T8G1,T4G1 2>&1 | tail -4
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $out/r02y_tests.log 2>&1
tail -4 $out/r02y_tests.log
timeout 600 python bench.py --workload ocs_batch --steps 100 --no-cpu-baseline --also none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('ocs_batch value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'hbm frac', round(r['frac'],3), 'mv_us', round(r['avg_launch_us'],1), 'share', round(r['share_of_step'],3), 'parity', d['parity']['ok'], d['parity']['parity_max_rel'])"
