#!/bin/bash
out=gpurun_out
mkdir -p $out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $out/r02f_tests.log 2>&1
tail -6 $out/r02f_tests.log
(time timeout 1500 python bench.py) > $out/r02f_bench.json 2> $out/r02f_bench.err
tail -4 $out/r02f_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench.json').read().strip().splitlines()[-1])
def brief(r):
    rf=r["roofline"]
    print(f"{r.get('name','HEAD'):10s} value {r['value']:.1f} ms/step {r['ms_per_step']:.3f} steps {r['steps']} e2e {r['e2e']['value']:.1f} | {rf['bound']} frac {rf['frac']:.3f} share {rf['share_of_step']:.2f} mv/ss {rf['matvecs_per_state_step']:.2f} mv_us {rf['avg_launch_us']:.1f} | parity {r['parity']['ok']} {r['parity']['parity_max_rel']:.1e} | launches {r['gpu_launches']}")
brief(d)
for r in d["workloads"]:
    if "error" in r: print(r); continue
    brief(r)
print(d["cpu_baseline"])
PY
RMB_E2E_TRACE=1 timeout 600 python bench.py --workload h2s --no-cpu-baseline --no-parity --also none --steps 5 2>&1 >/dev/null | grep "rmb e2e" | tail -2
RMB_E2E_TRACE=1 timeout 600 python bench.py --workload h2o --no-cpu-baseline --no-parity --also none 2>&1 >/dev/null | grep "rmb e2e" | tail -2
for l in 1 2 3; do RMB_LOOKAHEAD=$l timeout 600 python bench.py --workload h2o --no-cpu-baseline --no-parity --also none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('h2o lookahead $l', round(d['value']), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])"; done
