#!/bin/bash
out=gpurun_out
mkdir -p $out
(time timeout 900 python -m pytest tests -m gpu -x -q -k "dmma or h2s or g4 or wide_k or config4 or centrifuge or sub_batching") > $out/r02d_tests.log 2>&1
tail -5 $out/r02d_tests.log
timeout 600 python tools/matvec_probe.py h2s 64 3 2>&1 | tail -3 | tee $out/r02d_probe_h2s.log
timeout 600 python tools/matvec_probe.py asym 256 3 2>&1 | tail -3 | tee $out/r02d_probe_asym.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none -c 300 --csv --log-file $out/r02d_launches_h2s.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --also none > /dev/null 2>&1
timeout 900 ncu --metrics $M --clock-control none -c 500 --csv --log-file $out/r02d_launches_ocs.csv python bench.py --workload ocs_batch --steps 3 --warmup 3 --no-cpu-baseline --no-parity --also none > /dev/null 2>&1
timeout 900 ncu --metrics $M --clock-control none -c 500 --csv --log-file $out/r02d_launches_h2o.csv python bench.py --workload h2o --steps 3 --warmup 3 --no-cpu-baseline --no-parity --also none > /dev/null 2>&1
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py gemm > $out/r02d_racecheck_dmma.log 2>&1; tail -3 $out/r02d_racecheck_dmma.log
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py gemm > $out/r02d_memcheck_dmma.log 2>&1; tail -2 $out/r02d_memcheck_dmma.log
ls -la $out | grep r02d_
