#!/bin/bash
out=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $out/r02fz_tests.log 2>&1
tail -4 $out/r02fz_tests.log | head -2
for wl in ocs_align ocs_mixed; do RMB_LIB=build_variants/lib_ftrace.so timeout 200 python tools/fused_trace.py $wl 2>&1 | grep fused | head -4; done
timeout 900 python bench.py --workload h2o --also ocs_align,ocs_mixed,ocs_batch --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for r in [d]+d['workloads']:
    rf=r['roofline']; print(r.get('name','HEAD'), 'value', round(r['value'],1), 'ms/step', round(r['ms_per_step'],4), 'steps', r['steps'], 'e2e', round(r['e2e']['value'],1), 'mv/ss', round(rf['matvecs_per_state_step'],2), 'parity', r['parity']['ok'], r['parity']['parity_max_rel'], r['parity'].get('orders_equal'), 'multi', (r.get('multi_step_call') or {}).get('value'))"
