#!/bin/bash
out=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $out/r02q_tests.log 2>&1
tail -4 $out/r02q_tests.log
RMB_LIN_T=16 timeout 600 python tools/lin_soak.py 200 512 T8 2>&1 | tail -2
for T in 8 16; do
RMB_LIN_T=$T timeout 600 python tools/matvec_probe.py ocs_batch 8192 100 2>&1 | tail -2
RMB_LIN_T=$T timeout 600 python bench.py --workload ocs_batch --no-cpu-baseline --also none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('ocs_batch T=$T value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'hbm frac', round(r['frac'],3), 'mv_us', round(r['avg_launch_us'],1), 'share', round(r['share_of_step'],3), 'parity', d['parity']['ok'])"
done
timeout 600 python bench.py --workload ocs_align --steps 2000 --no-cpu-baseline --also none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ocs_align', round(d['value']), 'e2e', round(d['e2e']['value']), 'multi', d.get('multi_step_call'))"
