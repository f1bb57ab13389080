#!/bin/bash
out=gpurun_out
mkdir -p $out
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    r=d["roofline"]
    print(sys.argv[2], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "mv_us", round(r["avg_launch_us"],1), "parity", d["parity"] and d["parity"]["ok"])
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
for dk in 4 6 8 10; do
  RMB_GEMM_MIN_DK=$dk timeout 600 python bench.py --workload h2o --no-cpu-baseline --also none > $out/r02e_h2o_dk$dk.json 2>/dev/null
  show $out/r02e_h2o_dk$dk.json "h2o gemm_min_dk=$dk"
done
RMB_LIN_G1=1 timeout 600 python bench.py --workload ocs_batch --no-cpu-baseline --also none > $out/r02e_ocs_g1.json 2>/dev/null
show $out/r02e_ocs_g1.json "ocs_batch G1"
for c in 1 2 3 4 6; do
  RMB_E2E_TRACE=1 RMB_HOST_CHUNKS=$c timeout 600 python bench.py --workload h2s --no-cpu-baseline --no-parity --also none --steps 5 > $out/r02e_h2s_chunks$c.json 2> $out/r02e_h2s_chunks$c.err
  show $out/r02e_h2s_chunks$c.json "h2s chunks=$c"
  grep "rmb e2e" $out/r02e_h2s_chunks$c.err | tail -2
done
for c in 2 3 4; do
  RMB_E2E_TRACE=1 RMB_HOST_CHUNKS=$c timeout 600 python bench.py --workload h2o --no-cpu-baseline --no-parity --also none > $out/r02e_h2o_chunks$c.json 2> $out/r02e_h2o_chunks$c.err
  show $out/r02e_h2o_chunks$c.json "h2o chunks=$c"
  grep "rmb e2e" $out/r02e_h2o_chunks$c.err | tail -2
done
python tools/pcie_probe.py
