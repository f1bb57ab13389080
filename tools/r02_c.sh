#!/bin/bash
out=gpurun_out
mkdir -p $out
ls -la oracle/_ref | head
(time timeout 900 python -m pytest tests -m gpu -x -q -k "dmma or h2s or g4 or wide_k or config4 or centrifuge or sub_batching") > $out/r02c_tests.log 2>&1
tail -5 $out/r02c_tests.log
timeout 600 python tools/matvec_probe.py h2s 64 3 2>&1 | tail -3 | tee $out/r02c_probe_h2s.log
RMB_DMMA_TILES=1 timeout 600 python tools/matvec_probe.py h2s 64 3 2>&1 | tail -2 | tee -a $out/r02c_probe_h2s.log
timeout 600 python tools/matvec_probe.py asym 256 3 2>&1 | tail -3 | tee $out/r02c_probe_asym.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_matvec_dmma -s 1 -c 1 -o $out/r02c_dmma python tools/matvec_probe.py h2s 64 3 > $out/r02c_ncu.log 2>&1
tail -2 $out/r02c_ncu.log
(time timeout 1500 python bench.py --also asym,h2o) > $out/r02c_bench.json 2> $out/r02c_bench.err
tail -4 $out/r02c_bench.err
ls -la $out | grep r02c_
