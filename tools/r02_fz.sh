#!/bin/bash
out=gpurun_out
for wl in ocs_align ocs_mixed; do RMB_LIB=build_variants/lib_ftrace.so timeout 200 python tools/fused_trace.py $wl 2>&1 | grep fused | head -4; done
timeout 900 python -m pytest tests -m gpu -x -q -k "fused or ocs or linear or lanczos or zero or golden or g1 or g2 or g3" 2>&1 | tail -2
timeout 900 python bench.py --workload ocs_mixed --also ocs_align --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for r in [d]+d['workloads']:
    rf=r['roofline']; print(r.get('name','HEAD'), 'value', round(r['value'],1), 'ms/step', round(r['ms_per_step'],4), 'steps', r['steps'], 'e2e', round(r['e2e']['value'],1), 'parity', r['parity']['ok'], r['parity']['parity_max_rel'], r['parity'].get('orders_equal'), 'multi', (r.get('multi_step_call') or {}).get('value'))"
