#!/bin/bash
# One gpurun call that produces everything profiles/ needs for a round (about 10 GPU-minutes on one B200):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/round_profile.sh rNN'
# Outputs under gpurun_out/<tag>_*: GPU test log, the default bench line (all workloads), launch lists (duration + DRAM
# bytes per launch) of the three batched workloads and one `ncu --set full` capture of each matvec kernel and of the
# Lanczos recurrence kernel.  Summaries: tools/ncu_summary.py.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $out/${tag}_tests.log 2>&1
tail -3 $out/${tag}_tests.log
(time timeout 1800 python bench.py) > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -3 $out/${tag}_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for wl in h2s h2o ocs_batch; do
  timeout 900 ncu --metrics $M --clock-control none -c 400 --csv --log-file $out/${tag}_launches_$wl.csv \
      python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline --no-parity --also none > /dev/null 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_matvec_dmma -s 1 -c 1 -o $out/${tag}_dmma \
    python tools/matvec_probe.py h2s 64 3 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_matvec_tiled -s 1 -c 1 -o $out/${tag}_tiled \
    python tools/matvec_probe.py h2o 500 100 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_matvec_lin -s 1 -c 1 -o $out/${tag}_lin \
    python tools/matvec_probe.py ocs_batch 8192 100 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_recur_gram -s 3 -c 1 -o $out/${tag}_recur \
    python bench.py --workload h2s --steps 2 --warmup 3 --no-cpu-baseline --no-parity --also none > /dev/null 2>&1
python - <<PY
import json
d=json.loads(open("$out/${tag}_bench.json").read().strip().splitlines()[-1])
def brief(r):
    rf=r["roofline"]
    print(f"{r.get('name','HEAD'):10s} value {r['value']:.1f} ms/step {r['ms_per_step']:.3f} steps {r['steps']} e2e {r['e2e']['value']:.1f} | {rf['bound']} frac {rf['frac']:.3f} share {rf['share_of_step']:.2f} mv/ss {rf['matvecs_per_state_step']:.2f} mv_us {rf['avg_launch_us']:.1f} | parity {r['parity']['ok']} | launches {r['gpu_launches']} | cpu {r.get('cpu_baseline',{}).get('value')}")
brief(d)
for r in d["workloads"]:
    if "error" in r: print(r); continue
    brief(r)
print(d.get("cpu_baseline"))
PY
ls -la $out | grep ${tag}_
# device timelines of the bench step (CUPTI; shares and gaps only)
for wl in h2s h2o ocs_batch; do timeout 400 python tools/gpu_timeline.py $wl 4 2>&1 | grep -v "Warn\|_warn_once"; done > $out/${tag}_step_timeline.txt
