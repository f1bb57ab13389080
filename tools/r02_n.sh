#!/bin/bash
out=gpurun_out
M=gpu__time_duration.sum
for wl in ocs_mixed ocs_align; do
timeout 600 ncu --metrics $M --clock-control none -c 200 --csv --log-file $out/r02n_launches_$wl.csv python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-parity --also none > /dev/null 2>&1
python tools/ncu_summary.py launches $out/r02n_launches_$wl.csv | head -14
done
