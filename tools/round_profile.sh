#!/bin/bash
# One gpurun call that produces everything profiles/ needs for a round (about 4 GPU-minutes on one B200):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/round_profile.sh rNN'
# Outputs under gpurun_out/<tag>_*: GPU test log, bench lines of both workloads, launch lists (duration + DRAM bytes
# per launch) and one `ncu --set full` capture of each matvec kernel.  Summaries: tools/ncu_summary.py.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
(time python -m pytest tests -m gpu -x -q) > $out/${tag}_tests.log 2>&1
tail -3 $out/${tag}_tests.log
python bench.py > $out/${tag}_bench_h2o.json 2> $out/${tag}_bench_h2o.err
python bench.py --workload ocs --steps 40 > $out/${tag}_bench_ocs.json 2> $out/${tag}_bench_ocs.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -c 600 --csv --log-file $out/${tag}_launches_h2o.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics $M --clock-control none -c 400 --csv --log-file $out/${tag}_launches_ocs.csv \
    python bench.py --workload ocs --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_matvec_tiled -s 2 -c 1 -o $out/${tag}_tiled \
    python tools/matvec_probe.py h2o > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_matvec_lin -s 2 -c 1 -o $out/${tag}_lin \
    python tools/matvec_probe.py ocs > /dev/null 2>&1
python - <<PY
import json
for wl in ("h2o", "ocs"):
    try:
        d = json.load(open("$out/${tag}_bench_%s.json" % wl))
        r = d["roofline"]
        print(wl, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]),
              "matvec us", round(r["avg_launch_us"], 1), r["bound"], "frac", round(r["frac"], 3), "clocks", d["clocks"])
    except Exception as e:
        print(wl, "bench line missing:", e)
PY
ls -la $out | grep ${tag}_
