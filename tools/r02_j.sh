#!/bin/bash
out=gpurun_out
nvidia-smi topo -m 2>&1 | head -12
(time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --also h2o,ocs_mixed,ocs_batch,ocs_align) > $out/r02j_bench_n2.json 2> $out/r02j_bench_n2.err
tail -3 $out/r02j_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02j_bench_n2.json').read().strip().splitlines()[-1])
def brief(r):
    rf=r["roofline"]
    print(f"{r.get('name','HEAD'):10s} n{d['n_gpus']} {r['scaling']:7s} value {r['value']:.1f} ms/step {r['ms_per_step']:.3f} steps {r['steps']} e2e {r['e2e']['value']:.1f} | frac {rf['frac']:.3f} share {rf['share_of_step']:.2f} | parity {r['parity']['ok'] if r['parity'] else None}")
brief(d)
for r in d["workloads"]:
    if "error" in r: print(r); continue
    brief(r)
print(d["config"].get("cpu_affinity"))
PY
