#!/bin/bash
out=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $out/r02i_tests.log 2>&1
tail -4 $out/r02i_tests.log
for wl in h2s h2o ocs_batch ocs_align; do
RMB_E2E_TRACE=1 timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-parity --also none > $out/r02i_$wl.json 2> $out/r02i_$wl.err
python - <<PY
import json
d=json.load(open("$out/r02i_$wl.json"))
print("$wl value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"])
PY
grep "rmb e2e" $out/r02i_$wl.err | tail -1
done
for c in 3 4 8; do RMB_HOST_CHUNKS=$c timeout 600 python bench.py --workload h2s --no-cpu-baseline --no-parity --also none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('h2s chunks $c e2e', round(d['e2e']['value'],1))"; done
for c in 2 4 6; do RMB_HOST_CHUNKS=$c timeout 600 python bench.py --workload h2o --no-cpu-baseline --no-parity --also none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('h2o chunks $c e2e', round(d['e2e']['value'],1))"; done
