#!/bin/bash
out=gpurun_out
mkdir -p $out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $out/r02d_tests.log 2>&1
tail -15 $out/r02d_tests.log
timeout 600 python tools/matvec_probe.py h2s 64 3 2>&1 | tail -2 | tee $out/r02d_probe_h2s.log
timeout 600 python tools/matvec_probe.py asym 256 3 2>&1 | tail -2 | tee $out/r02d_probe_asym.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for g in 0 1; do
for wl in h2s ocs_batch h2o; do
RMB_GRAM=$g timeout 900 ncu --metrics $M --clock-control none -c 400 --csv --log-file $out/r02d_launches_${wl}_gram$g.csv python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-parity --also none > /dev/null 2>&1
done
done
for wl in h2s ocs_batch h2o; do
for g in 0 1; do
RMB_GRAM=$g timeout 900 python bench.py --workload $wl --no-cpu-baseline --also none 2>/dev/null > $out/r02d_bench_${wl}_gram$g.json
python - <<PY
import json
d=json.load(open("$out/r02d_bench_${wl}_gram$g.json"))
print("$wl gram$g value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "share", round(d["roofline"]["share_of_step"],3), "launches", d["gpu_launches"], "parity", d["parity"]["ok"], d["parity"]["parity_max_rel"])
PY
done
done
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py gemm > $out/r02d_racecheck_dmma.log 2>&1; tail -3 $out/r02d_racecheck_dmma.log
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py gemm > $out/r02d_memcheck_dmma.log 2>&1; tail -2 $out/r02d_memcheck_dmma.log
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py tiled > $out/r02d_memcheck_tiled.log 2>&1; tail -2 $out/r02d_memcheck_tiled.log
ls -la $out | grep r02d_
