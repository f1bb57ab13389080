#!/bin/bash
out=gpurun_out
mkdir -p $out
ls -la oracle/_ref oracle/_ref/richmol 2>&1 | head -12
(time timeout 900 python -m pytest tests -m gpu -x -q -k "dmma or h2s or g4 or wide_k or config4 or centrifuge") > $out/r02b_tests.log 2>&1
tail -15 $out/r02b_tests.log
timeout 600 python tools/matvec_probe.py h2s 64 3 2>&1 | tail -6 | tee $out/r02b_probe_h2s.log
timeout 600 python tools/matvec_probe.py asym 256 3 2>&1 | tail -6 | tee $out/r02b_probe_asym.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_matvec_dmma -s 1 -c 1 -o $out/r02b_dmma python tools/matvec_probe.py h2s 64 3 > $out/r02b_ncu.log 2>&1
tail -3 $out/r02b_ncu.log
timeout 600 python tools/lin_soak.py 300 8192 T8,T8G1,T4G1 2>&1 | tail -4 | tee $out/r02b_soak8192.log
ls -la $out | grep r02b_
