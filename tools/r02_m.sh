#!/bin/bash
out=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $out/r02m_tests.log 2>&1
tail -4 $out/r02m_tests.log
for wl in h2o h2s ocs_batch; do
for cr in 1 3; do
RMB_CORUN=$cr RMB_E2E_TRACE=1 timeout 600 python bench.py --workload $wl --no-cpu-baseline --also none > $out/r02m_${wl}_corun$cr.json 2> $out/r02m_${wl}_corun$cr.err
python - <<PY
import json
d=json.load(open("$out/r02m_${wl}_corun$cr.json"))
print("$wl corun=$cr value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], "parity", d["parity"]["ok"], d["parity"]["parity_max_rel"])
PY
grep "rmb e2e" $out/r02m_${wl}_corun$cr.err | tail -1
done
done
RMB_SPEC=0 timeout 600 python bench.py --workload h2o --no-cpu-baseline --no-parity --also none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('h2o spec=0', round(d['value']), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])"
